"""Launch single phases of the step kernel alone (for ncu): python tools/ncu_stage.py 2,14,29"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench as Bn
from jamie_b200.engine import Engine

phases = [int(p) for p in sys.argv[1].split(',')]
n = 200000
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
eng.train_steps(4)            # k_step launch 0: four whole steps
names = eng.phase_names()
for p in phases:              # then 4 launches per listed phase (3 warm-up + 1)
    us, _ = eng.bench_stage(p, 1)
    print(f'phase {p} {names[p]}: {us:.2f} us')
