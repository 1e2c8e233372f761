"""Benchmark of the JAMIE hot path on B200: train cells/sec (fwd + bwd + clip + Adam) of the coupled VAE.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json): N = 1 -> configs[1], synthetic scRNA+scATAC-like pair, 50k cells, post-PCA widths [512, 512],
output_dim 32, batch 512, 50 % partially matched P (diag mask -> 'hybrid' sampler), dropout 0.6, F = 0.
N > 1 -> configs[3], 1M cells sharded over the ranks, batch 512 per rank (weak scaling), one NCCL all-reduce of the
flat gradient buffer per step.  Data are synthetic standardised fp32 rows of the named shape; weights are the
reference's random init (torch seed 666).  A "step" is one optimizer step over one batch.

  value : cells/s, whole job, data and sampling plan resident in HBM, K CUDA-graph steps timed with CUDA events
  e2e   : cells/s through the C ABI with HOST data: per step the batch is gathered on the host into pinned memory,
          copied to the device inside jb_train_step_hostbatch, and the loss scalars are read back
  roofline     : the dominant kernel (largest TF32 GEMM stage) timed alone with CUDA events, vs the measured tensor peak
  step_roofline: compulsory HBM bytes of one step (fp32 param + Adam moments r/w + gathered inputs) / step time
  cpu_baseline : the numpy oracle (a port of the reference's algorithm) timed on this box's host cores
`--impl reference` times the oracle port alone (the reference is Python and /root/reference is absent on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS, LATENT, BATCH, DROPOUT = [512, 512], 32, 512, 0.6
# launch order of one training step (engine.cu: record_backward / record_update)
STEP_KERNELS = {
    31: ['k_gather(+step ctl)', 'k_corr_rowsum (side branch in the graph)', 'k_corr_build (side branch in the graph)',
         'gemm F1 enc D->2D', 'k_bn_fwd', 'gemm F2 enc 2D->D', 'k_bn_fwd', 'gemm F3 heads', 'k_reparam',
         'k_combine(+latent loss)', 'gemm F4 dec L->D', 'k_bn_fwd', 'gemm F5 dec D->2D', 'k_bn_fwd',
         'gemm F6 dec 2D->D', 'k_rec', 'gemm B6 dgrad', 'k_bn_bwd', 'gemm B5 dgrad', 'k_bn_bwd', 'gemm B4 dgrad',
         'k_latent_bwd_c', 'k_latent_bwd_z', 'k_latent_final', 'gemm B3 dgrad', 'k_bn_bwd', 'gemm B2 dgrad',
         'k_bn_bwd', 'gemm wgrad x12', 'k_gradnorm', 'k_adam'],
}
N_PARAMS = 4312194
# N > 1: 'single' = backward | one all-reduce of the flat gradient buffer | update (default: measured fastest at N = 2 and
# N = 8, profiles/README.md); 'overlap' = Engine.dp_step with bucket 0 all-reduced beside the encoder backward
DP_MODE = os.environ.get('JB_DP_MODE', 'single')


def load_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return {'hbm_gbs': pk['hbm_gbs'], 'bf16_tflops': pk['bf16_tflops'],
                'bf16_tflops_sustained': pk.get('bf16_tflops_sustained', pk['bf16_tflops']), 'source': 'measured'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_plan(n, steps, rng, corr_samples):
    """The reference's 'hybrid' sampler (jamie/jamie.py:559-573) for a partially matched diagonal prior."""
    idx0 = np.empty((steps, BATCH), np.int64)
    idx1 = np.empty((steps, BATCH), np.int64)
    for s in range(steps):
        k = min(int(np.sum(rng.random(BATCH) < .8)), 2)
        ci = rng.choice(2, k, replace=False)
        idx0[s] = np.concatenate([corr_samples[0][ci], rng.choice(n, BATCH - k, replace=False)])
        idx1[s] = np.concatenate([corr_samples[1][ci], rng.choice(n, BATCH - k, replace=False)])
    return idx0, idx1


def init_params():
    import torch
    from jamie_b200.model import edModelVar
    torch.manual_seed(666)
    m = edModelVar(DIMS, LATENT, dropout=DROPOUT)
    return m.packed_parameters(), m.packed_buffers()


def cpu_baseline(seconds=12.0, max_steps=200):
    """Oracle port of one reference optimizer step (sampling, gather, P block, fwd, losses, bwd, clip, Adam)."""
    from oracle import jamie_oracle as O
    rng = np.random.default_rng(0)
    n = 8192      # the reference's per-step cost does not depend on n
    data = [rng.standard_normal((n, d)).astype(np.float32) for d in DIMS]
    m = (rng.random(n) < 0.5).astype(np.float32)
    nz = np.flatnonzero(m)[:2]
    params, _ = init_params()
    orc = O.OracleModel(DIMS, LATENT, dropout=DROPOUT, params=params)
    cs = np.stack([nz, nz], 1)

    def one(step):
        i0, i1 = make_plan(n, 1, rng, cs)
        i0, i1 = i0[0], i1[0]
        x = [data[0][i0], data[1][i1]]
        Pb = (i0[:, None] == i1[None, :]).astype(np.float32) * m[i0][:, None]
        s = Pb.sum(1)
        s[s == 0] = 1
        corr = Pb / s[:, None]
        eps = [rng.standard_normal((BATCH, LATENT)).astype(np.float32) for _ in range(2)]
        widths = [1024, 512, 1024, 512, 512, 1024, 512, 1024]
        masks = [(rng.random((BATCH, w)) >= DROPOUT).astype(np.uint8) for w in widths]
        orc.train_step(x, corr, np.zeros((BATCH, BATCH), np.float32), eps, masks, 0.5)

    for s in range(2):
        one(s)
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps and time.perf_counter() - t0 < seconds:
        one(steps)
        steps += 1
    dt = time.perf_counter() - t0
    return {'value': BATCH * steps / dt, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': f'{steps} optimizer steps of batch {BATCH} at widths {DIMS}, output_dim {LATENT} '
                      f'(numpy fp32 oracle, BLAS on all host cores), {dt:.1f} s', 'ms_per_step': 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import jamie_oracle as O  # noqa: F401
    t_budget = max(5.0, min(120.0, 0.15 * (args.steps + args.warmup)))
    cb = cpu_baseline(seconds=t_budget, max_steps=max(args.steps, 3))
    line = {
        'impl': 'reference', 'metric': 'train cells/sec (fwd+bwd+Adam)', 'value': cb['value'], 'unit': 'cells/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'oracle port of the reference step on the host cores: the reference is pure Python/torch and cannot '
                'travel to the GPU box',
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        return t.get(kernel_key)
    except Exception:
        return None


def predict_leg(eng, torch, peaks, stream, rows_dev=1 << 20, rows_host=1 << 18):
    """modal_predict (BASELINE metric 2): modality 0 -> 1 on pre-transformed [N, 512] fp32 rows.
    value: rows resident in HBM, CUDA events; e2e: pinned host buffers in and out through jb_predict."""
    g = torch.Generator(device='cuda').manual_seed(2)
    x = torch.randn((rows_dev, DIMS[0]), generator=g, device='cuda', dtype=torch.float32)
    out = torch.empty((rows_dev, DIMS[1]), device='cuda', dtype=torch.float32)
    l0 = eng.launch_count()
    eng.predict_into(0, 1, x.data_ptr(), 1 << 16, DIMS[0], out.data_ptr(), DIMS[1], 1, stream)   # warm-up, folds BN
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.predict_into(0, 1, x.data_ptr(), rows_dev, DIMS[0], out.data_ptr(), DIMS[1], 1, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert bool(torch.isfinite(out[::4097]).all())
    del x, out
    hx = torch.randn((rows_host, DIMS[0]), dtype=torch.float32).pin_memory()
    ho = torch.empty((rows_host, DIMS[1]), dtype=torch.float32).pin_memory()
    eng.predict_into(0, 1, hx.data_ptr(), rows_host, DIMS[0], ho.data_ptr(), DIMS[1], 0, stream)
    t0 = time.perf_counter()
    eng.predict_into(0, 1, hx.data_ptr(), rows_host, DIMS[0], ho.data_ptr(), DIMS[1], 0, stream)
    dt = time.perf_counter() - t0
    flops_row = 2.0 * (2 * DIMS[0] * DIMS[0] * 2 + DIMS[0] * LATENT + LATENT * DIMS[1] + 2 * DIMS[1] * DIMS[1] * 2)
    v = rows_dev / (ms * 1e-3)
    return {'metric': 'modal_predict cells/sec', 'value': v, 'unit': 'cells/s', 'rows': rows_dev, 'ms': ms,
            'e2e': {'value': rows_host / dt, 'unit': 'cells/s', 'rows': rows_host,
                    'h2d_bytes': rows_host * DIMS[0] * 4, 'd2h_bytes': rows_host * DIMS[1] * 4},
            'roofline': {'bound': 'tensor', 'achieved': v * flops_row / 1e12, 'peak': peaks['bf16_tflops_sustained'],
                         'unit': 'TFLOP/s', 'frac': v * flops_row / 1e12 / peaks['bf16_tflops_sustained'],
                         'flops_per_row': flops_row,
                         'note': 'BatchNorm folded, single-pass TF32 (hardware rate = half the bf16 peak in the denominator)'},
            'gpu_launches': eng.launch_count() - l0}


def workload_config(n_gpus):
    if n_gpus == 1:
        wl = 'BASELINE configs[1]: synthetic 50k-cell pair, post-PCA widths [512,512], output_dim 32, batch 512, ' \
             '50% partially matched diagonal P (hybrid sampler), dropout 0.6, F=0'
        n = 50_000
    else:
        wl = f'BASELINE configs[3]: synthetic 1M-cell pair sharded over {n_gpus} ranks, widths [512,512], output_dim 32, ' \
             f'batch 512 per rank, 50% partially matched diagonal P, dropout 0.6, F=0, one flat-gradient all-reduce/step'
        n = 1_000_000
    return {'workload': wl, 'cells': n, 'widths': DIMS, 'output_dim': LATENT, 'batch_per_rank': BATCH,
            'parallelism': f'dp{n_gpus}', 'l2_policy': 'parameter + Adam state (69 MB) and gathered rows are re-read '
            'every step by design; inputs are gathered from a dataset larger than L2 (N>1) / re-sampled rows (N=1)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-predict', action='store_true')
    ap.add_argument('--predict-only', action='store_true', help='modal_predict leg alone (tuning runs)')
    ap.add_argument('--profile', action='store_true', help='device-resident steps only (for ncu): no e2e / stage / CPU legs, no JSON line')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    # stdout carries exactly one JSON line: anything libraries print on the way (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        # 17 MB all-reduce per step: measured on 8 B200 375.7 us/step with NVLS off vs 392.5 us/step with NCCL's default
        os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from jamie_b200.engine import Engine
    peaks = load_peaks()
    K, W = args.steps, max(args.warmup, 3)

    n_total = workload_config(args.gpus)['cells']
    n = n_total // world
    g = torch.Generator(device='cuda').manual_seed(1234 + rank)
    data = [torch.randn((n, d), generator=g, device='cuda', dtype=torch.float32) for d in DIMS]
    rng = np.random.default_rng(100 + rank)
    mask = (rng.random(n) < 0.5).astype(np.float32)
    nz = np.flatnonzero(mask)[:2]
    cs = np.stack([nz, nz], 1)
    params, bufs = init_params()
    eng = Engine(DIMS, LATENT, BATCH, DROPOUT, seed=666 * 1000003 + rank, device=local, world_size=world)
    eng.set_params(params)
    eng.set_bn_stats(bufs)
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(2):
        eng.set_dataset(i, data[i], stream)
    eng.set_prior_diag(mask)
    eng.set_f_dense(None)
    idx0, idx1 = make_plan(n, W + K, rng, cs)
    eng.upload_plan(idx0, idx1, np.full(W + K, 0.5), stream)
    gt = eng.grad_tensor() if world > 1 else None
    buckets = [eng.grad_bucket_tensor(0), eng.grad_bucket_tensor(1)] if world > 1 else None
    if args.predict_only:
        p = predict_leg(eng, torch, peaks, stream)
        emit({k: p[k] for k in ('value', 'ms', 'e2e', 'gpu_launches')})
        eng.close()
        return

    def run_steps(k):
        if world == 1:
            eng.train_steps(k, stream)
        else:
            if DP_MODE == 'single':
                for _ in range(k):  # backward | one all-reduce of the flat gradient buffer | update
                    eng.step_backward(stream)
                    dist.all_reduce(gt)
                    eng.step_update(stream)
            else:                   # backward part 0 | all-reduce(bucket 0) beside backward part 1 | all-reduce(bucket 1) | update
                for _ in range(k):
                    eng.dp_step(dist, buckets, stream, overlap=DP_MODE == 'overlap')

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    run_steps(W)
    sync_all()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        run_steps(K)
        e1.record()
        torch.cuda.synchronize()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    losses = eng.read_losses(W + K, stream)
    assert np.all(np.isfinite(losses[:, :6])), 'non-finite losses in the timed region'
    value = BATCH * K * world / (ms * 1e-3)

    if args.profile:
        print(f'profile run: {ms / K * 1e3:.1f} us/step', flush=True)
        eng.close()
        return
    prof = None
    if world == 1:
        eng.upload_plan(idx0[:1], idx1[:1], np.full(1, 0.5), stream)
        us = eng.profile_step(20, stream)
        names = STEP_KERNELS.get(len(us), [f'launch{k}' for k in range(len(us))])
        prof = {'sum_us': float(us.sum()), 'launches': [[n, round(float(u), 2)] for n, u in zip(names, us)],
                'note': 'eager launches with a CUDA event between consecutive kernels (warm L2, no graph, no PDL): '
                        'shares of the step, not absolutes'}

    # ---------------- end to end with host-resident data
    Ke = max(20, min(K, 300))
    host = [d.cpu().pin_memory() for d in data] if n <= 200_000 else \
        [d[:200_000].cpu().pin_memory() for d in data]
    nh = host[0].shape[0]
    hb = [[torch.empty((BATCH, d), dtype=torch.float32).pin_memory() for d in DIMS] for _ in range(2)]   # two host slots
    i0e, i1e = make_plan(nh, Ke + 5, rng, cs if nz.max() < nh else np.stack([np.flatnonzero(mask[:nh])[:2]] * 2, 1))
    ti = [torch.from_numpy(i0e), torch.from_numpy(i1e)]

    def e2e_run(first, count):
        """The reference's loop shape (jamie/jamie.py:549-742) with host-resident data: gather the batch rows on the host,
        hand them to the engine, read the step's losses back. N = 1: the engine's asynchronous host-batch API keeps two
        steps in flight, so the host gather of batch k + 1 overlaps step k; N > 1: backward, all-reduce, update."""
        out, pending = None, 0
        for s in range(first, first + count):
            b = hb[s & 1]
            for i in range(2):
                torch.index_select(host[i], 0, ti[i][s], out=b[i])      # the reference's dataset[i][random_batch[i]]
            if world == 1:
                eng.hostbatch_submit(b[0].data_ptr(), b[1].data_ptr(), i0e[s], i1e[s], 0.5, stream)
                pending += 1
                if pending == 2:
                    out = eng.hostbatch_wait()
                    pending -= 1
            else:
                eng.step_backward_hostbatch(b[0].data_ptr(), b[1].data_ptr(), i0e[s], i1e[s], 0.5, stream)
                dist.all_reduce(gt)
                eng.step_update(stream)
                out = eng.read_losses(2, stream)[s & 1]
        while pending:
            out = eng.hostbatch_wait()
            pending -= 1
        return out

    e2e_run(0, 5)
    sync_all()
    t0 = time.perf_counter()
    out = e2e_run(5, Ke)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    assert np.isfinite(out[4])
    e2e_value = BATCH * Ke * world / dt
    h2d = sum(BATCH * d * 4 for d in DIMS) + 2 * BATCH * 4 + 4
    d2h = 8 * 4

    pred = predict_leg(eng, torch, peaks, stream) if not args.no_predict else None
    if pred is not None and world > 1:      # rows/s of the whole job: every rank imputes its own shard, no collective
        t = torch.tensor([pred['ms'], pred['rows'] / pred['e2e']['value']], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pred['value'] = world * pred['rows'] / (float(t[0]) * 1e-3)
        pred['e2e']['value'] = world * pred['e2e']['rows'] / float(t[1])
    line = None
    if rank == 0:
        # ---------------- dominant kernel alone (CUDA events on the launching stream)
        stages = []
        for st in range(12):
            us, fl = eng.bench_stage(st, 200, stream)
            stages.append((us, fl, st))
        tot_gemm_us = sum(s[0] for s in stages)
        us, fl, st = max(stages)
        tf = fl / (us * 1e-6) / 1e12
        step_us = ms * 1e3 / K
        alg_bytes = 6 * 4 * N_PARAMS + 4 * BATCH * sum(DIMS)
        roof = {'bound': 'tensor', 'kernel': f'gemm_tf32_grouped_kernel (stage {st})', 'achieved': tf,
                'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': tf / peaks['bf16_tflops'],
                'traffic': None, 'us_per_launch': us, 'flops_per_launch': fl, 'peak_source': peaks['source'],
                'note': 'kernel computes in TF32 (hardware rate = half the bf16 peak used as denominator); '
                        f'all 12 GEMM stages alone sum to {tot_gemm_us:.1f} us of the {step_us:.1f} us step'}
        sroof = {'bound': 'hbm', 'achieved': alg_bytes / (step_us * 1e-6) / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                 'frac': alg_bytes / (step_us * 1e-6) / 1e9 / peaks['hbm_gbs'], 'bytes_per_step': alg_bytes,
                 'note': 'algorithmic bytes of a whole step = r/w of fp32 params + Adam m, v (6*4*4312194) + gathered '
                         'inputs; SURVEY.md section 8d: the step roofline is 16.1 us'}
        cb = None
        roof['traffic'] = ncu_traffic('gemm_stage_%d' % st)
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline()
        line = {
            'metric': 'train cells/sec (fwd+bwd+Adam)', 'value': value, 'unit': 'cells/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'tf32', 'data': 'synthetic',
            'config': workload_config(world),
            'e2e': {'value': e2e_value, 'unit': 'cells/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': Ke, 'note': 'per step: host gather of the batch rows into pinned memory, H2D of rows + cell ids, '
                                         'step graph, D2H of the 8 loss scalars; jb_hostbatch_submit / jb_hostbatch_wait keep '
                                         'two steps in flight (N = 1)'},
            'gpu_launches': int(launches), 'launches_per_step': launches / K,
            'roofline': roof, 'step_roofline': sroof, 'step_profile': prof,
            'gemm_stages_us': [round(s[0], 2) for s in stages], 'modal_predict': pred, 'cpu_baseline': cb, 'clocks': clocks.summary(),
            'final_losses': {k: float(v) for k, v in zip(['KL', 'Rec', 'CosSim', 'F', 'total', 'grad_norm'], losses[-1])},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == '__main__':
    main()
