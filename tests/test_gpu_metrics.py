"""GPU evaluation metrics (jb_metric_*) against the numpy oracle of the same definitions
(oracle/metrics_oracle.py, itself pinned against the reference's sklearn calls in tests/test_host_logic.py) and against the
reference's own metric values on its MMD-MA run (tests/golden/mmdma.npz holds the embeddings' metrics, not the
embeddings, so that part lives in test_gpu_mmdma.py)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


def _pair(n, L, seed, noise=0.4):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, L)).astype(np.float32)
    b = (a + noise * rng.normal(size=a.shape)).astype(np.float32)
    return a, b, rng


@pytest.mark.parametrize('n,L', [(60, 4), (300, 32), (1000, 17), (2500, 32), (65, 70)])
def test_foscttm_counts_are_exact(n, L):
    from jamie_b200 import evaluation as E
    a, b, _ = _pair(n, L, n + L, noise=0.35 * np.sqrt(L))     # noisy enough that some cells ARE closer than the match
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        got = E.test_closer([a, b])
    assert buf.getvalue().startswith('foscttm: ')              # the reference prints this line
    want = MO.test_closer([a, b], verbose=False)
    assert got == want, (got, want)                            # integer counts over float64 distances: exact
    assert E.test_closer([a, a], verbose=False) == 0.0
    assert 0.0 < got < 0.5


def test_foscttm_matches_the_distance_metric_hook():
    """the reference's signature: a callable on the concatenated embeddings gives the same number"""
    from sklearn.metrics import pairwise_distances
    from jamie_b200 import evaluation as E
    a, b, _ = _pair(200, 8, 3)
    assert E.test_closer([a, b], verbose=False) == E.test_closer(
        [a, b], distance_metric=lambda x: pairwise_distances(x, metric='euclidean'), verbose=False)


@pytest.mark.parametrize('n0,n1,L,ncls,k', [(60, 45, 4, 3, None), (60, 45, 4, 3, 1), (500, 700, 32, 7, 5), (300, 300, 16, 4, 41),
                                             (1200, 900, 32, 11, None), (40, 40, 8, 2, 40)])
def test_label_transfer_matches_oracle(n0, n1, L, ncls, k):
    from jamie_b200 import evaluation as E
    rng = np.random.default_rng(n0 + n1)
    y0 = rng.integers(0, ncls, size=n0)
    y1 = rng.integers(1, ncls + 1, size=n1)                    # a different label set, as the reference allows
    centers = rng.normal(size=(ncls + 1, L)) * 1.5
    emb0 = (centers[y0] + rng.normal(size=(n0, L))).astype(np.float32)
    emb1 = (centers[y1] + rng.normal(size=(n1, L))).astype(np.float32)
    got, k_got = E.label_transfer_accuracy([emb0, emb1], [y0, y1], k=k, return_k=True)
    want, k_want = MO.label_transfer_accuracy([emb0, emb1], [y0, y1], k=k, return_k=True)
    assert k_got == k_want
    assert got == want, (got, want)


def test_label_transfer_ties_go_to_the_lowest_index_and_class():
    """duplicated reference rows (replicated cells) tie exactly: a stable argsort admits the lowest indices first, and a
    tied vote goes to the lowest class"""
    from jamie_b200 import evaluation as E
    rng = np.random.default_rng(9)
    base = rng.normal(size=(30, 6)).astype(np.float32)
    emb1 = np.concatenate([base, base, base])                  # every distance appears three times
    y1 = np.concatenate([np.zeros(30, int), np.ones(30, int), np.full(30, 2)])
    emb0 = (base + 0.01 * rng.normal(size=base.shape)).astype(np.float32)
    y0 = np.zeros(30, int)
    # query i's three nearest rows are the copies of base[i] at rows i, 30 + i, 60 + i (classes 0, 1, 2), exactly tied:
    # k = 1 must take row i (class 0); k = 2 and 3 tie the vote -> class 0; k = 4, 6 add class-0 rows of the next triple
    for k in (1, 2, 3, 4, 6):
        assert E.label_transfer_accuracy([emb0, emb1], [y0, y1], k=k) == 1.0, k
    assert E.test_LabelTA([emb0, emb1], [y0, y1], verbose=False) == E.label_transfer_accuracy([emb0, emb1], [y0, y1], k=5)


def test_feature_pearson_matches_oracle():
    from jamie_b200 import evaluation as E
    rng = np.random.default_rng(2)
    x = rng.normal(size=(700, 130)).astype(np.float32) * 3 + 5
    y = (0.6 * x + rng.normal(size=x.shape)).astype(np.float32)
    y[:, 7] = 1.0                                              # a constant measured feature: nan, skipped by the mean
    with np.errstate(all='ignore'):
        want = MO.imputation_correlation(x, y)
    got = E.imputation_correlation(x, y)
    keep = np.isfinite(want)
    assert not np.isfinite(got[7]) and not keep[7]
    np.testing.assert_allclose(got[keep], want[keep], rtol=0, atol=1e-9)
    assert E.mean_feature_r(x, y) == pytest.approx(MO.mean_feature_r(x, y), abs=1e-9)
    np.testing.assert_allclose(E.imputation_correlation(x * 2 + 1, x), 1.0, atol=1e-9)


def test_gpu_metrics_reproduce_the_reference_values():
    """the GPU metrics on the embeddings of tests/golden/metrics.npz == what the reference's own functions returned"""
    import json
    import os
    from jamie_b200 import evaluation as E
    from tests.golden_util import GOLDEN_DIR
    G = np.load(os.path.join(GOLDEN_DIR, 'metrics.npz'))
    for c, rec in enumerate(json.loads(str(G['meta']))):
        e0, e1, y0, y1 = (G[f'c{c}/{k}'] for k in ('e0', 'e1', 'y0', 'y1'))
        if 'foscttm' in rec:
            assert E.test_closer([e0, e1], verbose=False) == rec['foscttm']
        for k in (1, 5, 17):
            assert E.test_LabelTA([e0, e1], [y0, y1], k=k, verbose=False) == rec[f'lta_k{k}']
        acc, kdef = E.label_transfer_accuracy([e0, e1], [y0, y1], k=None, return_k=True)
        assert kdef == rec['k_default'] and acc == rec['lta_default']
        want = G[f'c{c}/r']
        keep = np.isfinite(want)
        np.testing.assert_allclose(E.imputation_correlation(G[f'c{c}/x'], G[f'c{c}/y'])[keep], want[keep], atol=1e-6)
