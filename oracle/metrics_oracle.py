"""TEST INFRASTRUCTURE ONLY.  CPU (numpy, float64) restatement of the reference's evaluation metrics: the checker of
``jamie_b200/evaluation.py`` (which computes them on the GPU through jb_metric_*).  Only ``tests/`` may import it.

Reference: FOSCTTM ``test_closer`` jamie/evaluation.py:65-85 (class method jamie/jamie.py:892-913), label-transfer
accuracy ``test_LabelTA`` jamie/evaluation.py:114-132 (class method jamie/jamie.py:943-961: k = None default), per-feature
imputation correlation jamie/evaluation.py:491-513 (sklearn ``r_regression``).  Pinned by ``tests/test_host_logic.py::
test_metrics`` against the sklearn calls the reference itself makes (pairwise_distances, KNeighborsClassifier) and by the
reference's own metric values stored in ``tests/golden/mmdma.npz``.

Distances are |a|^2 + |b|^2 - 2 a.b in float64 over row chunks (the reference materialises the (2n) x (2n) sklearn
distance matrix).
"""
import numpy as np

_CHUNK = 2048


def _sqdist_rows(a, b, lo, hi):
    """Squared euclidean distances of rows a[lo:hi] to every row of b (float64)."""
    aa = (a[lo:hi] ** 2).sum(1)[:, None]
    bb = (b ** 2).sum(1)[None, :]
    d = aa + bb - 2.0 * (a[lo:hi] @ b.T)
    np.maximum(d, 0.0, out=d)
    return d


def test_closer(integrated_data, distance_metric=None, verbose=True):
    """Fraction of samples closer than the true match: raw_count / (2 size^2), both directions (lower is better)."""
    assert len(integrated_data) == 2, 'Two datasets are supported for FOSCTTM'
    a = np.asarray(integrated_data[0], np.float64)
    b = np.asarray(integrated_data[1], np.float64)
    size = a.shape[0]
    raw_count_closer = 0
    if distance_metric is not None:   # the reference's signature: a callable on the concatenated embeddings
        distances = distance_metric(np.concatenate([a, b], axis=0))
        for i in range(size):
            local = distances[i][size:]
            raw_count_closer += int(np.sum(local < local[i]))
            local = distances[size + i][:size]
            raw_count_closer += int(np.sum(local < local[i]))
    else:
        for lo in range(0, size, _CHUNK):
            hi = min(size, lo + _CHUNK)
            idx = np.arange(lo, hi)
            d = _sqdist_rows(a, b, lo, hi)          # A -> B
            raw_count_closer += int((d < d[np.arange(hi - lo), idx][:, None]).sum())
            d = _sqdist_rows(b, a, lo, hi)          # B -> A
            raw_count_closer += int((d < d[np.arange(hi - lo), idx][:, None]).sum())
    foscttm = raw_count_closer / (2 * size ** 2)
    if verbose:
        print(f'foscttm: {foscttm}')
    return foscttm


foscttm = test_closer


def default_k(datatype):
    """20 % of the average class size (jamie/jamie.py:946-950)."""
    total_size = min(*[len(d) for d in datatype])
    num_classes = len(np.unique(np.concatenate(datatype)).flatten())
    return int(.2 * total_size / num_classes)


def test_LabelTA(integrated_data, datatype, k=5, return_k=False, verbose=True):
    """kNN classifier (uniform votes, euclidean) fitted on modality 1's embedding and labels, scored on modality 0's."""
    if k is None:
        k = default_k(datatype)
    emb0 = np.asarray(integrated_data[0], np.float64)
    emb1 = np.asarray(integrated_data[1], np.float64)
    y0 = np.asarray(datatype[0]).ravel()
    y1 = np.asarray(datatype[1]).ravel()
    classes, y1c = np.unique(y1, return_inverse=True)   # sklearn: classes sorted, ties -> the lowest class
    pred = np.empty(emb0.shape[0], dtype=classes.dtype)
    for lo in range(0, emb0.shape[0], _CHUNK):
        hi = min(emb0.shape[0], lo + _CHUNK)
        d = _sqdist_rows(emb0, emb1, lo, hi)
        nn = np.argsort(d, axis=1, kind='stable')[:, :k]
        votes = np.zeros((hi - lo, len(classes)), np.int64)
        np.add.at(votes, (np.arange(hi - lo)[:, None], y1c[nn]), 1)
        pred[lo:hi] = classes[np.argmax(votes, axis=1)]
    acc = float(np.sum(pred == y0)) / len(y0)
    if verbose:
        print(f'label transfer accuracy: {acc}')
    if return_k:
        return acc, k
    return acc


def label_transfer_accuracy(integrated_data, datatype, k=None, return_k=False):
    """The class method's form (jamie/jamie.py:943-961): k defaults to 20 % of the average class size, nothing printed."""
    return test_LabelTA(integrated_data, datatype, k=k, return_k=return_k, verbose=False)


def imputation_correlation(imputed, measured):
    """Per-feature Pearson r between imputed and measured values (constant features give nan and are skipped)."""
    x = np.asarray(imputed, np.float64)
    y = np.asarray(measured, np.float64)
    xc = x - x.mean(0)
    yc = y - y.mean(0)
    den = np.sqrt((xc ** 2).sum(0) * (yc ** 2).sum(0))
    with np.errstate(all='ignore'):
        r = (xc * yc).sum(0) / den
    return r


def mean_feature_r(imputed, measured):
    """Mean over the non-constant measured features, as ``_plot_correlation`` aggregates (jamie/evaluation.py:491-513)."""
    y = np.asarray(measured)
    keep = np.array([len(np.unique(col)) > 1 for col in y.T])
    return float(np.nanmean(imputation_correlation(imputed, measured)[keep]))
