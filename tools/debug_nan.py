"""Find the first non-finite step of the bench workload and dump tap statistics around it."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench as Bn
from jamie_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
chunks = [int(c) for c in sys.argv[2].split(',')] if len(sys.argv) > 2 else [10]
total = int(sys.argv[3]) if len(sys.argv) > 3 else 400
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
idx0, idx1 = Bn.make_plan(n, total, rng, cs)
eng.upload_plan(idx0, idx1, np.full(total, 0.5))
B, L, D = Bn.BATCH, Bn.LATENT, Bn.DIMS
names = ['x', 'y1_', 'h1_', 'y2_', 'h2_', 'mulv', 'z', 'c', 'g1_', 'g2_', 'xhat', 'dxhat', 'dg2_', 'dy4_', 'dg1_', 'dy3_', 'dc', 'dmulv', 'dh2_', 'dy2_', 'dh1_', 'dy1_']
widths = {'x': 1, 'y1_': 2, 'h1_': 2, 'y2_': 1, 'h2_': 1, 'mulv': -2, 'z': -1, 'c': -1, 'g1_': 1, 'g2_': 2, 'xhat': 1, 'dxhat': 1, 'dg2_': 2, 'dy4_': 2,
          'dg1_': 1, 'dy3_': 1, 'dc': -1, 'dmulv': -2, 'dh2_': 1, 'dy2_': 1, 'dh1_': 2, 'dy1_': 2}
done = 0
ci = 0
while done < total:
    chunk = min(chunks[min(ci, len(chunks) - 1)], total - done)
    ci += 1
    eng.train_steps(chunk)
    done += chunk
    ls = eng.read_losses(done)
    bad = np.flatnonzero(~np.isfinite(ls[:, :6]).all(axis=1))
    last = ls[done - 1]
    print(f'steps {done}: last losses {last[:6]}', flush=True)
    if len(bad):
        print('first non-finite step', bad[0], ls[bad[0]], 'previous', ls[max(0, bad[0] - 1)])
    if chunk == 1 or done % 50 == 0 or len(bad):
        for nm in names:
            for i in range(2):
                w = widths[nm]
                cols = (D[i] * w) if w > 0 else (L * -w)
                t = eng.debug_read(f'{nm}{i}', (B, cols))
                if not np.isfinite(t).all():
                    w_ = np.argwhere(~np.isfinite(t))
                    print(f'   {nm}{i} non-finite count {len(w_)} first {w_[:6].tolist()} rows {np.unique(w_[:,0])[:8]} cols {np.unique(w_[:,1])[:8]}')
                print(f'   {nm}{i:<2d} absmax {np.nanmax(np.abs(t)):.4g}  rms {np.sqrt((t.astype(np.float64) ** 2).mean()):.4g}  finite {np.isfinite(t).all()}')
    if len(bad):
        break
