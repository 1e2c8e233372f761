"""Per-phase timeline of the persistent step kernel on the bench workload: total, longest / mean CTA work, barrier tail."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench as Bn
from jamie_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 48
torch.cuda.set_device(0)
g = torch.Generator(device='cuda').manual_seed(1234)
data = [torch.randn((n, d), generator=g, device='cuda', dtype=torch.float32) for d in Bn.DIMS]
rng = np.random.default_rng(100)
mask = (rng.random(n) < 0.5).astype(np.float32)
nz = np.flatnonzero(mask)[:2]
cs = np.stack([nz, nz], 1)
params, bufs = Bn.init_params()
eng = Engine(Bn.DIMS, Bn.LATENT, Bn.BATCH, Bn.DROPOUT, seed=666 * 1000003)
eng.set_params(params); eng.set_bn_stats(bufs)
for i in range(2):
    eng.set_dataset(i, data[i])
eng.set_prior_diag(mask); eng.set_f_dense(None)
idx0, idx1 = Bn.make_plan(n, 64, rng, cs)
eng.upload_plan(idx0, idx1, np.full(64, 0.5))
eng.train_steps(8)
eng.upload_plan(idx0, idx1, np.full(64, 0.5))
us = eng.profile_step(iters)
det = eng.profile_detail()
print(f'{"phase":24s} {"total":>8s} {"work max":>9s} {"work avg":>9s} {"tail":>7s}')
for nm, d in zip(eng.phase_names(), det):
    print(f'{nm:24s} {d[0]:8.2f} {d[1]:9.2f} {d[2]:9.2f} {d[3]:7.2f}')
print(f'{"sum":24s} {det[:, 0].sum():8.2f} {det[:, 1].sum():9.2f} {det[:, 2].sum():9.2f} {det[:, 3].sum():7.2f}')
print('GEMM phases, CTA 0, us after the phase began: producer enters | first TMA issued | last TMA issued | first stage landed | last stage landed | accumulator read | tile stored')
gn = [nm for nm in eng.phase_names() if nm.startswith(('gemm', 'dgrad', 'wgrad'))]
for nm, r in zip(gn, eng.gemm_stamps):
    print(f'{nm:24s} ' + ' '.join(f'{v:7.2f}' for v in r))
if os.environ.get('JB_STAGES'):
    base, _ = eng.bench_stage(-1, 200)
    print(f'empty launch {base:.2f} us')
    for p, nm in enumerate(eng.phase_names()):
        us1, _ = eng.bench_stage(p, 200)
        print(f'stage {nm:24s} alone {us1:8.2f} us/launch   minus empty {us1 - base:8.2f}')
