"""Per-tensor gradient errors of one injected step (GPU engine vs numpy oracle). Usage: grad_check.py D0 D1 L B p"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import jamie_oracle as O
from tests import parity_util as U
from jamie_b200.engine import Engine

a = [float(x) for x in sys.argv[1:]]
D0, D1, L, B = int(a[0]), int(a[1]), int(a[2]), int(a[3]); p = a[4]
dims = [D0, D1]; n = 2 * B
data = U.synth_pair(n, dims, seed=1); params = U.torch_like_init(dims, L, seed=2)
rng = np.random.default_rng(5)
m = (rng.random(n) < 0.5).astype(np.float32)
eng = Engine(dims, L, B, p); eng.set_params(params)
for i in range(2): eng.set_dataset(i, data[i])
eng.set_prior_diag(m); eng.set_f_dense(None)
orc = O.OracleModel(dims, L, dropout=p, params=params)
i0 = rng.choice(n, B, replace=False); i1 = np.concatenate([i0[:B // 2], rng.choice(n, B - B // 2, replace=False)])
eng.upload_plan(i0[None], i1[None], np.array([0.37]))
eps, masks = U.draw_randomness(B, dims, L, p, seed=11)
eng.inject(eps, masks); eng.train_steps(1)
ls = eng.read_losses(1)[0]
corr = O.corr_block(np.diag(m), i0, i1)
ols, og, otot, fw = orc.train_step([data[0][i0], data[1][i1]], corr, np.zeros((B, B), np.float32), eps, masks, 0.37)
print('env', {k: v for k, v in os.environ.items() if k.startswith('JB_')}, 'losses', ls[:6], [float(v) for v in ols], otot)
for key in ['dxhat', 'dg2_', 'dy4_', 'dg1_', 'dy3_', 'dc', 'dmulv', 'dh2_', 'dy2_', 'dh1_', 'dy1_']:
    for i in range(2):
        want = orc.last_bwd_taps[f'{key}{i}']
        got = eng.debug_read(f'{key}{i}', want.shape)
        cm = np.abs(want.mean(0)).mean() / np.abs(want).mean()
        print(f'  tap {key}{i:d}  rel {U.rel(got, want):.3e}   colmean/abs {cm:.2e}  rel(colsum) {U.rel(got.sum(0), want.sum(0)):.3e}')
for (nm, _), g in zip(orc.spec, eng.get_grads()):
    print(f'  {nm:24s} rel {U.rel(g, og[nm]):.3e}  |g| {np.linalg.norm(g):.3e} |ref| {np.linalg.norm(og[nm]):.3e}')
