"""The acceptance metrics the reference's notebooks report, with the reference's own definitions (same denominators,
same default k, same printed lines), computed on the GPU: the O(n^2 L) distance work of FOSCTTM and
of the kNN label transfer, and the per-feature correlation of an imputed matrix, run as CUDA kernels behind the C ABI
(``jb_metric_foscttm``, ``jb_metric_knn_vote``, ``jb_metric_feature_pearson``: csrc/metrics.cuh).

Reference: FOSCTTM ``test_closer`` jamie/evaluation.py:65-85 (class method jamie/jamie.py:892-913), label-transfer
accuracy ``test_LabelTA`` jamie/evaluation.py:114-132 (class method jamie/jamie.py:943-961: k = None default), per-feature
imputation correlation jamie/evaluation.py:491-513 (sklearn ``r_regression``).

There is no host fallback: without the CUDA library these functions raise.  The only host-side distance computation left
is the reference's ``distance_metric=`` hook of ``test_closer`` (a user callable on the concatenated embeddings).
"""
import ctypes as C

import numpy as np

from . import _lib


def _f32(x):
    return np.ascontiguousarray(np.asarray(x), np.float32)


def _device(device):
    if device is not None:
        return int(device)
    import torch
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def test_closer(integrated_data, distance_metric=None, verbose=True, device=None):
    """Fraction of samples closer than the true match: raw_count / (2 size^2), both directions (lower is better)."""
    assert len(integrated_data) == 2, 'Two datasets are supported for FOSCTTM'
    size = np.asarray(integrated_data[0]).shape[0]
    if distance_metric is not None:   # the reference's signature: a callable on the concatenated embeddings
        a = np.asarray(integrated_data[0], np.float64)
        b = np.asarray(integrated_data[1], np.float64)
        distances = distance_metric(np.concatenate([a, b], axis=0))
        raw_count_closer = 0
        for i in range(size):
            local = distances[i][size:]
            raw_count_closer += int(np.sum(local < local[i]))
            local = distances[size + i][:size]
            raw_count_closer += int(np.sum(local < local[i]))
    else:
        a, b = _f32(integrated_data[0]), _f32(integrated_data[1])
        assert a.shape == b.shape and a.ndim == 2, 'FOSCTTM needs two [n, L] embeddings of matched rows'
        cnt = C.c_ulonglong(0)
        lib = _lib.load()
        _lib.check(lib.jb_metric_foscttm(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1],
                                         _device(device), C.byref(cnt)))
        raw_count_closer = int(cnt.value)
    foscttm = raw_count_closer / (2 * size ** 2)
    if verbose:
        print(f'foscttm: {foscttm}')
    return foscttm


foscttm = test_closer   # name used by the tests of this repository


def default_k(datatype):
    """20 % of the average class size (jamie/jamie.py:946-950)."""
    total_size = min(*[len(d) for d in datatype])
    num_classes = len(np.unique(np.concatenate(datatype)).flatten())
    return int(.2 * total_size / num_classes)


def test_LabelTA(integrated_data, datatype, k=5, return_k=False, verbose=True, device=None):
    """kNN classifier (uniform votes, euclidean) fitted on modality 1's embedding and labels, scored on modality 0's."""
    if k is None:
        k = default_k(datatype)
    emb0, emb1 = _f32(integrated_data[0]), _f32(integrated_data[1])
    y0 = np.asarray(datatype[0]).ravel()
    y1 = np.asarray(datatype[1]).ravel()
    classes, y1c = np.unique(y1, return_inverse=True)   # sklearn: classes sorted, ties -> the lowest class
    y1c = np.ascontiguousarray(y1c, np.int32)
    pred = np.empty(emb0.shape[0], np.int32)
    lib = _lib.load()
    _lib.check(lib.jb_metric_knn_vote(emb0.ctypes.data_as(C.c_void_p), emb0.shape[0], emb1.ctypes.data_as(C.c_void_p),
                                      y1c.ctypes.data_as(C.c_void_p), emb1.shape[0], emb0.shape[1], int(k), len(classes),
                                      _device(device), pred.ctypes.data_as(C.c_void_p)))
    acc = float(np.sum(classes[pred] == y0)) / len(y0)
    if verbose:
        print(f'label transfer accuracy: {acc}')
    if return_k:
        return acc, k
    return acc


def label_transfer_accuracy(integrated_data, datatype, k=None, return_k=False):
    """The class method's form (jamie/jamie.py:943-961): k defaults to 20 % of the average class size, nothing printed."""
    return test_LabelTA(integrated_data, datatype, k=k, return_k=return_k, verbose=False)


def imputation_correlation(imputed, measured, device=None):
    """Per-feature Pearson r between imputed and measured values (constant features give nan and are skipped)."""
    x, y = _f32(imputed), _f32(measured)
    assert x.shape == y.shape and x.ndim == 2
    r = np.empty(x.shape[1], np.float64)
    lib = _lib.load()
    _lib.check(lib.jb_metric_feature_pearson(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), x.shape[0], x.shape[1],
                                             _device(device), r.ctypes.data_as(C.c_void_p)))
    return r


def mean_feature_r(imputed, measured):
    """Mean over the non-constant measured features, as ``_plot_correlation`` aggregates (jamie/evaluation.py:491-513)."""
    y = np.asarray(measured)
    keep = np.array([len(np.unique(col)) > 1 for col in y.T])
    return float(np.nanmean(imputation_correlation(imputed, measured)[keep]))
