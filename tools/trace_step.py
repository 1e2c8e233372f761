"""In-graph timeline of one training step: CUPTI kernel records (through torch.profiler) of the step graph as it really
runs (CUDA graph + programmatic dependent launch + the forked P/F branch), which neither ncu (serialises, cold cache) nor
jb_profile_step (eager, an event between launches) can show.

  python tools/trace_step.py [--steps 6] [--out gpurun_out/trace_step.json]

Prints one row per kernel of the middle step: start offset inside the step, duration, gap to the previous kernel's end.
A profiler is attached, so these are NOT bench numbers: use them for shares and gaps only.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'trace_step.json'))
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench as Bn
    from jamie_b200.engine import Engine
    torch.cuda.set_device(0)
    n = 50_000
    g = torch.Generator(device='cuda').manual_seed(1234)
    data = [torch.randn((n, d), generator=g, device='cuda', dtype=torch.float32) for d in Bn.DIMS]
    rng = np.random.default_rng(100)
    mask = (rng.random(n) < 0.5).astype(np.float32)
    nz = np.flatnonzero(mask)[:2]
    cs = np.stack([nz, nz], 1)
    params, bufs = Bn.init_params()
    eng = Engine(Bn.DIMS, Bn.LATENT, Bn.BATCH, Bn.DROPOUT, seed=1)
    eng.set_params(params)
    eng.set_bn_stats(bufs)
    for i in range(2):
        eng.set_dataset(i, data[i])
    eng.set_prior_diag(mask)
    eng.set_f_dense(None)
    total = 50 + args.steps
    idx0, idx1 = Bn.make_plan(n, total, rng, cs)
    eng.upload_plan(idx0, idx1, np.full(total, 0.5))
    eng.train_steps(50)
    torch.cuda.synchronize()
    tmp = args.out + '.chrome.json'
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        eng.train_steps(args.steps)
        torch.cuda.synchronize()
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))['traceEvents']
    ks = sorted((e for e in ev if e.get('cat') == 'kernel'), key=lambda e: e['ts'])
    os.remove(tmp)
    if not ks:
        print('no kernel records captured')
        return
    # a step starts at its k_gather
    starts = [i for i, e in enumerate(ks) if 'k_gather' in e['name']]
    if len(starts) < 3:
        print('fewer than 3 steps captured')
        return
    a = starts[len(starts) // 2]
    b = starts[len(starts) // 2 + 1]
    mid = ks[a:b]
    per = b - a
    t0 = mid[0]['ts']
    rows, prev_end = [], t0
    for e in mid:
        rows.append({'name': e['name'][:60], 'start_us': round(e['ts'] - t0, 2), 'dur_us': round(e['dur'], 2),
                     'gap_us': round(e['ts'] - prev_end, 2), 'grid': e.get('args', {}).get('grid'),
                     'stream': e.get('args', {}).get('stream')})
        prev_end = max(prev_end, e['ts'] + e['dur'])
    span = prev_end - t0
    step_all = (ks[-1]['ts'] + ks[-1]['dur'] - ks[0]['ts']) / args.steps
    json.dump({'kernels_per_step': per, 'step_span_us': span, 'avg_step_us_under_profiler': step_all, 'rows': rows},
              open(args.out, 'w'), indent=1)
    print(f'{per} kernels/step, middle step spans {span:.1f} us, {step_all:.1f} us/step under the profiler')
    for r in rows:
        print(f"{r['start_us']:8.2f} {r['dur_us']:7.2f} gap {r['gap_us']:6.2f}  {r['name']}")
    eng.close()


if __name__ == '__main__':
    main()
