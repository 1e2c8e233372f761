"""Bring-up: one injected training step on the GPU engine vs the numpy oracle, printing per-tensor errors.
Usage (GPU box): python tools/bringup_step.py [D0 D1 L B p]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import jamie_oracle as O  # noqa: E402
from tests import parity_util as U  # noqa: E402
from jamie_b200.engine import Engine  # noqa: E402


def main():
    a = [float(x) for x in sys.argv[1:]]
    D0, D1, L, B = (int(a[0]), int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 else (512, 512, 32, 512)
    p = a[4] if len(a) >= 5 else 0.6
    fmode = int(a[5]) if len(a) >= 6 else 0
    dims = [D0, D1]
    n = 4 * B
    data = U.synth_pair(n, dims, seed=1)
    params = U.torch_like_init(dims, L, seed=2)
    rng = np.random.default_rng(3)
    m = (rng.random(n) < 0.5).astype(np.float32)
    Fd = None
    lw = [1, 1, 1, 1]
    pf = 1.0
    if fmode:
        Fd = (rng.random((n, n)) * (rng.random((n, n)) < 0.05)).astype(np.float32)
        lw = [1, 2, 0.5, 3]
        pf = 0.7
    eng = Engine(dims, L, B, p, loss_weights=lw, pf_ratio=pf)
    eng.set_params(params)
    for i in range(2):
        eng.set_dataset(i, data[i])
    eng.set_prior_diag(m)
    eng.set_f_dense(Fd)
    orc = O.OracleModel(dims, L, dropout=p, params=params)
    nsteps = 3
    idx0 = np.stack([rng.choice(n, B, replace=False) for _ in range(nsteps)])
    idx1 = idx0.copy()
    idx1[:, B // 2:] = np.stack([rng.choice(n, B - B // 2, replace=False) for _ in range(nsteps)])
    anneal = np.array([O.kl_anneal(e, 4, 10) for e in range(nsteps)])
    eng.upload_plan(idx0, idx1, anneal)
    P = np.diag(m)
    Fm = np.zeros((n, n), np.float32) if Fd is None else Fd
    for s in range(nsteps):
        eps, masks = U.draw_randomness(B, dims, L, p, seed=10 + s)
        eng.inject(eps, masks)
        t0 = time.time()
        eng.train_steps(1)
        losses = eng.read_losses(s + 1)[s]
        t1 = time.time()
        x = [data[i][[idx0, idx1][i][s]] for i in range(2)]
        Pb = O.corr_block(P, idx0[s], idx1[s])
        Fb = O.corr_block(Fm, idx0[s], idx1[s])
        corr = (np.float32(pf) * Pb + np.float32(1 - pf) * Fb).astype(np.float32)
        theta_before = [q.copy() for q in orc.param_list()]
        ls, grads, tot, fw = orc.train_step(x, corr, Fb, eps, masks, anneal[s], lw)
        print(f'--- step {s}  (gpu call {1e3 * (t1 - t0):.2f} ms)  nnz(corr) {int((corr != 0).sum())}')
        print('  corr      ', U.rel(eng.debug_read('corr', (B, B)), corr))
        ot = U.oracle_taps(fw, orc)
        for name in ['x', 'h1_', 'h2_', 'mulv', 'z', 'c', 'g1_', 'g2_', 'xhat']:
            for i in range(2):
                key = f'{name}{i}'
                got = eng.debug_read(key, ot[key].shape)
                print(f'  {key:10s} rel {U.rel(got, ot[key]):.3e}')
        print('  losses gpu', losses[:6], ' oracle', [float(v) for v in ls], 'norm', tot)
        gg = eng.get_grads()
        gworst = 0
        for (nm, _), g_ in zip(orc.spec, gg):
            if nm in U.PRE_BN_BIAS:
                continue
            r_ = U.rel(g_, grads[nm])
            gworst = max(gworst, r_)
            if r_ > 3e-4:
                print(f'  grad {nm:24s} rel {r_:.3e}')
        print(f'  worst per-tensor gradient error {gworst:.3e}')
        after = eng.get_params()
        worst = 0
        for (nm, _), ga, oa, tb in zip(orc.spec, after, orc.param_list(), theta_before):
            d_g = ga - tb
            d_o = oa - tb
            e = float(np.abs(d_g - d_o).max())
            worst = max(worst, e)
            if nm in U.PRE_BN_BIAS:
                continue
            print(f'  dtheta {nm:24s} max|gpu-oracle| {e:.3e}   rel(update) {U.rel(d_g, d_o):.3e}')
        print('  worst abs param diff', worst)
        # keep both sides in lock-step for the next step
        eng.set_params(orc.param_list())
    print('launches', eng.launch_count())


if __name__ == '__main__':
    main()
