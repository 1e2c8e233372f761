"""Pins the numpy oracle (oracle/jamie_oracle.py) against fixtures recorded from the real reference.

Every reference-side quantity of each optimizer step is replayed: numpy sampler draws, P/F blocks, forward tensors,
the 4 losses, all 45 gradient tensors, the clip norm, the post-Adam parameters, BatchNorm buffers, and at the end the
embeddings, ``transform_one`` and ``modal_predict`` outputs.  CPU only.
"""
import numpy as np
import pytest

from oracle import jamie_oracle as O
from tests.golden_util import CASES, Golden

# pre-BatchNorm Linear biases: mathematically zero gradient, the reference holds fp32 noise
PRE_BN_BIAS = {f'{k}.{i}.{l}.bias' for k in ('encoders', 'decoders') for i in (0, 1) for l in (0, 4)}


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('name', CASES)
def test_replay(name):
    G = Golden(name)
    kw = G.kw
    dims = G.meta['col']
    L = kw['output_dim']
    B = G.meta['batch_size']
    rows = G.meta['n']
    model = O.OracleModel(dims, L, dropout=G.meta['dropout'], params=G.init_params(), buffers=G.buffers('init'),
                          dist_method=kw.get('dist_method', 'euclidean'))
    assert [n for n, _ in model.spec] == G.param_names
    P = G.P_dense(); Fm = G.F_dense()
    method = O.sampling_method(P)
    assert method == G.meta['sampling_method']
    corr_samples = np.argwhere(P > 0) if method == 'hybrid' else None
    data = [G[f'pre{i}'].astype(np.float32) for i in range(2)]
    pf = kw.get('PF_Ratio') or 1
    lw = kw.get('loss_weights')
    len_dl = max(int(max(rows) / kw['batch_size']), 1)
    np.random.seed(42)
    for s in range(G.n_steps):
        epoch = s // len_dl
        idx = O.sample_batch(method, rows, dims, B, corr_samples)
        rec = G.choices(s)
        if method == 'diag':
            assert np.array_equal(idx[0], rec[0])
        elif method == 'zeros':
            assert np.array_equal(idx[0], rec[0]) and np.array_equal(idx[1], rec[1])
        else:
            assert np.array_equal(idx[0][len(rec[0]):], rec[1]) and np.array_equal(idx[1][len(rec[0]):], rec[2])
        x = [data[i][idx[i]] for i in range(2)]
        for i in range(2):
            np.testing.assert_array_equal(x[i], G[f's{s}/x{i}'])
        Pb = O.corr_block(P, idx[0], idx[1]); Fb = O.corr_block(Fm, idx[0], idx[1])
        corr = (np.float32(pf) * Pb + np.float32(1 - pf) * Fb).astype(np.float32)
        np.testing.assert_allclose(corr, G[f's{s}/corr'], rtol=1e-6, atol=1e-7)
        anneal = O.kl_anneal(epoch, kw['min_epochs'], kw['epoch_DNN'])
        ls, grads, tot, fw = model.train_step(x, corr, Fb, G.eps(s), G.masks(s), anneal, lw)
        for i in range(2):
            assert rel(fw['mu'][i], G[f's{s}/mu{i}']) < 2e-5
            assert rel(fw['z'][i], G[f's{s}/z{i}']) < 2e-5
            assert rel(fw['c'][i], G[f's{s}/c{i}']) < 2e-5
            assert rel(fw['xhat'][i], G[f's{s}/xhat{i}']) < 2e-5
        assert rel(fw['lv'][1], G[f's{s}/logvars']) < 2e-5
        if (s + 1) % len_dl == 0:   # loss_history holds the last batch of each epoch, weighted (jamie.py:752-761)
            for k, nm in enumerate(['KL', 'Rec', 'CosSim', 'F']):
                want = G[f'loss_history/{nm}'][epoch]
                got = float(ls[k]) * (lw[k] if lw else 1)
                # CosSim goes through cdist's matmul route in the reference: absolute floor
                assert abs(got - want) <= 2e-5 * abs(want) + (2e-5 if nm == 'CosSim' else 1e-7), (nm, got, want)
        gref = G.grads(s)
        noise = set(PRE_BN_BIAS)
        if np.abs(corr).sum() == 0:
            noise.add('sigma')       # corr == 0 => c_i == z_i: d(loss)/d(sigma) is mathematically zero
        for (n, _), gr in zip(model.spec, gref):
            if n in noise:
                assert np.abs(grads[n]).max() < 1e-5 and np.abs(gr).max() < 1e-5
            else:
                assert rel(grads[n], gr) < 2e-4, (s, n, rel(grads[n], gr))
        assert abs(tot - float(G[f's{s}/total_norm'])) < 1e-4 * tot
        for (n, _), pr in zip(model.spec, G.params_after(s)):
            tol = 2.5e-3 if n in noise else 3e-5       # Adam normalises noise gradients to +-lr steps (App. B-18)
            assert np.abs(model.params[n] - pr).max() < tol, (s, n, np.abs(model.params[n] - pr).max())
        for k, v in G.buffers(f's{s}').items():
            if k.endswith('num_batches_tracked'):
                assert int(model.buffers[k]) == int(v)
            else:
                np.testing.assert_allclose(model.buffers[k], v, rtol=2e-5, atol=2e-6)
        # continue from the reference's exact state so that steps are pinned independently
        for (n, _), pr in zip(model.spec, G.params_after(s)):
            model.params[n] = pr.astype(np.float32).copy()
    # end-of-run eval paths
    for i in range(2):
        assert rel(model.encode_mu(data[i], i), G[f'tone{i}']) < 2e-5
        assert rel(model.encode_mu(data[i], i), G[f'emb{i}']) < 2e-5


@pytest.mark.parametrize('name', ['diag_drop', 'pca', 'rep_F'])
def test_modal_predict_and_preclass(name):
    G = Golden(name)
    kw = G.kw
    dims = G.meta['col']
    model = O.OracleModel(dims, kw['output_dim'], dropout=G.meta['dropout'], params=G.params_after(G.n_steps - 1),
                          buffers=G.buffers(f's{G.n_steps - 1}'))
    pres = []
    for i in range(2):
        raw = G[f'data{i}']
        if kw.get('pca_dim') is not None:
            comp, mean = G[f'pca_components{i}'], G[f'pca_mean{i}']
            sample = (raw - mean) @ comp.T
            pres.append(O.OraclePre(sample, comp, mean, axis=None))
        else:
            pres.append(O.OraclePre(raw, axis=0))
        assert rel(pres[i].transform(raw), G[f'pre{i}']) < 1e-9
    for i in range(2):
        to = 1 - i
        dec = model.impute(pres[i].transform(G[f'data{i}']).astype(np.float32), i, to)
        got = pres[to].inverse_transform(dec)
        assert rel(got, G[f'pred{i}']) < 2e-5


def test_metrics_oracle_against_the_reference_functions():
    """oracle/metrics_oracle.py against values the unmodified reference's own evaluation functions produced
    (tests/golden/metrics.npz, written by tests/golden/make_metrics.py: jamie.evaluation.test_closer / test_LabelTA, the
    JAMIE class methods with their default k, sklearn r_regression per feature)."""
    import json
    import os
    from oracle import metrics_oracle as MO
    from tests.golden_util import GOLDEN_DIR
    G = np.load(os.path.join(GOLDEN_DIR, 'metrics.npz'))
    meta = json.loads(str(G['meta']))
    for c, rec in enumerate(meta):
        e0, e1, y0, y1 = (G[f'c{c}/{k}'] for k in ('e0', 'e1', 'y0', 'y1'))
        if 'foscttm' in rec:
            assert MO.test_closer([e0, e1], verbose=False) == rec['foscttm'] == rec['foscttm_method']
        for k in (1, 5, 17):
            assert MO.test_LabelTA([e0, e1], [y0, y1], k=k, verbose=False) == rec[f'lta_k{k}']
        acc, kdef = MO.label_transfer_accuracy([e0, e1], [y0, y1], k=None, return_k=True)
        assert kdef == rec['k_default'] and acc == rec['lta_default']
        with np.errstate(all='ignore'):
            r = MO.imputation_correlation(G[f'c{c}/x'], G[f'c{c}/y'])
        want = G[f'c{c}/r']
        keep = np.isfinite(want)
        assert not keep[5] and keep.sum() == 11
        np.testing.assert_allclose(r[keep], want[keep], atol=1e-6)      # r_regression works in the inputs' float32
