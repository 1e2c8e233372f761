"""How large may the latent get before a step leaves the finite range? One step with fc_vars.1.bias = b for growing b
(|z| ~ exp(b / 2)), engine vs the fp32 oracle: which gradients / taps are non-finite on either side."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import jamie_oracle as O
from tests import parity_util as U
from jamie_b200.engine import Engine

dims, L, B, p = ([512, 512], 32, 512, 0.6) if os.environ.get('JB_RANGE_HEADLINE') else ([96, 64], 8, 64, 0.3)
n = 2 * B
for bias in [float(x) for x in (sys.argv[1:] or [24, 30, 36, 40, 44])]:
    rng = np.random.default_rng(15)
    data = U.synth_pair(n, dims, seed=31)
    params = U.torch_like_init(dims, L, seed=32)
    names = [nm for nm, _ in O.param_spec(dims, L)]
    params[names.index('fc_vars.1.bias')][:] = bias
    eng = Engine(dims, L, B, p)
    eng.set_params(params)
    for i in range(2):
        eng.set_dataset(i, data[i])
    m = (rng.random(n) < 0.5).astype(np.float32)
    eng.set_prior_diag(m); eng.set_f_dense(None)
    orc = O.OracleModel(dims, L, dropout=p, params=params)
    i0 = rng.choice(n, B, replace=False)
    i1 = np.concatenate([i0[:B // 2], rng.choice(n, B - B // 2, replace=False)])
    eng.upload_plan(i0[None], i1[None], np.array([0.5]))
    eps, masks = U.draw_randomness(B, dims, L, p, seed=33)
    eng.inject(eps, masks)
    eng.train_steps(1)
    ls = eng.read_losses(1)[0]
    Pb = O.corr_block(np.diag(m), i0, i1)
    with np.errstate(all='ignore'):
        ols, ograds, otot, fw = orc.train_step([data[0][i0], data[1][i1]], Pb.astype(np.float32), np.zeros_like(Pb), eps, masks, 0.5, None)
    g = eng.get_grads()
    bad_e = [nm for (nm, _), x in zip(eng.spec, g) if not np.isfinite(x).all()]
    bad_o = [nm for nm, _ in orc.spec if not np.isfinite(ograds[nm]).all()]
    worst = max((U.rel(x, ograds[nm]) for (nm, _), x in zip(eng.spec, g) if nm not in U.PRE_BN_BIAS and np.isfinite(x).all() and np.isfinite(ograds[nm]).all()), default=float('nan'))
    print(f'bias {bias}: max|z1| {np.abs(fw["z"][1]).max():.3g}  losses eng {ls[:6]}  oracle {[float(v) for v in ols]} |g| {otot:.4g}')
    print(f'   non-finite gradient tensors: engine {len(bad_e)} {bad_e[:4]}  oracle {len(bad_o)} {bad_o[:4]}; worst rel err of the finite ones {worst:.2e}')
    for key in ('z1', 'c1', 'c0', 'g1_0', 'g1_1', 'xhat0', 'xhat1'):
        shape = U.oracle_taps(fw, orc)[key].shape
        t = eng.debug_read(key, shape)
        if not np.isfinite(t).all():
            print(f'   tap {key}: non-finite in the engine')
    bshapes = {'dxhat': dims, 'dg2_': [2 * d for d in dims], 'dy4_': [2 * d for d in dims], 'dg1_': dims, 'dy3_': dims, 'dc': [L, L],
               'dmulv': [2 * L, 2 * L], 'dh2_': dims, 'dy2_': dims, 'dh1_': [2 * d for d in dims], 'dy1_': [2 * d for d in dims]}
    for key, w in bshapes.items():
        for i in range(2):
            t = eng.debug_read(f'{key}{i}', (B, w[i]))
            if not np.isfinite(t).all():
                print(f'   backward tap {key}{i}: non-finite in the engine ({int((~np.isfinite(t)).sum())} of {t.size}), finite max {np.nanmax(np.abs(np.where(np.isfinite(t), t, 0))):.3g}')
    eng.close()
