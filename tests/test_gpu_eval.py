"""GPU parity of the eval-mode paths (transform_one / modal_predict / PCA projection) against the oracle and the
reference fixtures, including ragged sizes and both pointer kinds."""
import numpy as np
import pytest

from oracle import jamie_oracle as O
from tests import parity_util as U
from tests.golden_util import Golden

pytestmark = pytest.mark.gpu
EVAL_RTOL = 2e-3


def _trained_state(dims, L, seed):
    params = U.torch_like_init(dims, L, seed=seed)
    rng = np.random.default_rng(seed + 1)
    bufs = O.init_buffers(dims)
    for k in bufs:
        if k.endswith('running_mean'):
            bufs[k] = rng.normal(size=bufs[k].shape).astype(np.float32) * 0.3
        elif k.endswith('running_var'):
            bufs[k] = (0.5 + rng.random(bufs[k].shape)).astype(np.float32)
    # non-trivial BN affine
    spec = O.param_spec(dims, L)
    for j, (n, shp) in enumerate(spec):
        if n != 'sigma' and (n.split('.')[-2] in ('1', '5')) and not n.startswith('fc_'):
            params[j] = (1 + 0.2 * rng.normal(size=shp)).astype(np.float32) if n.endswith('weight') else \
                (0.1 * rng.normal(size=shp)).astype(np.float32)
    return params, bufs


@pytest.mark.parametrize('dims,L,n', [([512, 512], 32, 2 * 128 * 148 + 8192 + 37), ([512, 39], 32, 300), ([40, 12], 6, 1), ([72, 40], 8, 129)])
def test_encode_predict_vs_oracle(dims, L, n):
    from jamie_b200.engine import Engine
    import torch
    params, bufs = _trained_state(dims, L, 3)
    eng = Engine(dims, L, 64, 0.5)
    eng.set_params(params)
    eng.set_bn_stats(bufs)
    orc = O.OracleModel(dims, L, dropout=0.5, params=params, buffers=bufs)
    data = U.synth_pair(n, dims, seed=9)
    for i in range(2):
        want_mu = orc.encode_mu(data[i], i)
        got = eng.encode(i, data[i])
        assert got.shape == want_mu.shape
        assert U.rel(got, want_mu) < EVAL_RTOL
        want_x = orc.impute(data[i], i, 1 - i)
        got_h = eng.predict(i, 1 - i, data[i])
        assert U.rel(got_h, want_x) < EVAL_RTOL
        xd = torch.from_numpy(data[i]).cuda()
        got_d = eng.predict(i, 1 - i, xd).cpu().numpy()
        np.testing.assert_array_equal(got_d, got_h)          # host-streamed and device-resident paths agree exactly
        # non-contiguous rows (pitch > width) on the device
        wide = torch.zeros((n, dims[i] + 5), device='cuda')
        wide[:, :dims[i]] = xd
        got_w = eng.predict(i, 1 - i, wide[:, :dims[i]]).cpu().numpy()
        np.testing.assert_allclose(got_w, got_h, rtol=1e-6, atol=1e-6)
    eng.close()


@pytest.mark.parametrize('name', ['diag_drop', 'pca', 'rep_F'])
def test_reference_fixture_eval(name):
    """transform_one and impute on the reference's final state reproduce the reference's recorded outputs."""
    from jamie_b200.engine import Engine
    G = Golden(name)
    dims, L = G.meta['col'], G.kw['output_dim']
    eng = Engine(dims, L, G.meta['batch_size'], G.meta['dropout'])
    eng.set_params(G.params_after(G.n_steps - 1))
    eng.set_bn_stats(G.buffers(f's{G.n_steps - 1}'))
    for i in range(2):
        pre = G[f'pre{i}'].astype(np.float32)
        assert U.rel(eng.encode(i, pre), G[f'tone{i}']) < EVAL_RTOL
        assert U.rel(eng.encode(i, pre), G[f'emb{i}']) < EVAL_RTOL
    eng.close()


@pytest.mark.parametrize('n,d,k', [(700, 1302, 64), (33, 100, 16), (2100, 39, 39)])
def test_pca_project_inverse(n, d, k):
    import ctypes as C
    from jamie_b200.engine import Engine, _ptr
    from jamie_b200 import _lib
    rng = np.random.default_rng(0)
    X = (rng.normal(size=(n, d)) * (1 + rng.random(d)) + rng.normal(size=d)).astype(np.float32)
    mean = X.mean(0).astype(np.float32)
    q, _ = np.linalg.qr(rng.normal(size=(d, k)))
    comp = np.ascontiguousarray(q.T.astype(np.float32))
    m, s = 0.3, 1.7
    eng = Engine([8, 8], 4, 8, 0.0)
    out = np.empty((n, k), np.float32)
    _lib.check(eng.lib.jb_pca_project(eng.h, _ptr(X), n, d, _ptr(comp), _ptr(mean), k, m, s, _ptr(out), 0, None))
    want = ((X.astype(np.float64) - mean) @ comp.T.astype(np.float64) - m) / s
    assert U.rel(out, want) < 2e-6          # 3xTF32 split: fp32-level accuracy
    back = np.empty((n, d), np.float32)
    _lib.check(eng.lib.jb_pca_inverse(eng.h, _ptr(out), n, k, _ptr(comp), _ptr(mean), d, m, s, _ptr(back), 0, None))
    want_b = (out.astype(np.float64) * s + m) @ comp.astype(np.float64) + mean
    assert U.rel(back, want_b) < 2e-6
    eng.close()
