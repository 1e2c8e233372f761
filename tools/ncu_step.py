"""Two launches of the whole-step kernel for ncu: launch 0 = 4 warm-up steps, launch 1 = NSTEPS whole steps (the one to
capture: `ncu -k regex:k_step -s 1 -c 1`). Prints the event-timed duration of launch 1 when run without a profiler."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench as Bn
from jamie_b200.engine import Engine

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
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
idx0, idx1 = Bn.make_plan(n, 4 + nsteps, rng, cs)
eng.upload_plan(idx0, idx1, np.full(4 + nsteps, 0.5))
eng.train_steps(4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.train_steps(nsteps); e1.record(); torch.cuda.synchronize()
print(f'k_step launch of {nsteps} steps: {e0.elapsed_time(e1) * 1e3 / nsteps:.1f} us/step')
