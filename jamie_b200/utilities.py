"""Host-side helpers that must exist under the reference's import paths for checkpoint compatibility
(``jamie.utilities.preclass`` / ``jamie.utilities.identity`` are pickled by reference inside saved models) plus the
lap timer whose labels the reference prints with ``debug=True``.

Reference: jamie/utilities.py:48-50 (identity), :61-132 (time_logger), :654-678 (preclass).
"""
import warnings
from time import perf_counter

import numpy as np


def identity(x):
    """Identity preprocessing (a module-level function so that it pickles)."""
    return x


class time_logger():
    """Lap timer keyed by label: ``log(label)`` records the time since the previous call, ``aggregate()`` prints the
    per-label means and their total (same output format as the reference's)."""

    def __init__(self, discard_first_sample=False, record=True, verbose=False, memory_usage=False):
        self.discard_first_sample = discard_first_sample
        self.record = record
        self.verbose = verbose
        self.memory_usage = memory_usage
        self.history = {}
        self.history_mem = {}
        if memory_usage:
            import tracemalloc
            tracemalloc.start()
        self.start_time = perf_counter()

    def log(self, str=''):
        if not (self.verbose or self.record):
            return
        now = perf_counter()
        lap = now - self.start_time
        if self.record:
            self.history.setdefault(str, []).append(lap)
        if self.verbose:
            print(f'{str}: {lap}')
        if self.memory_usage:
            import tracemalloc
            self.history_mem.setdefault(str, []).append(tracemalloc.get_traced_memory())
            tracemalloc.stop()
            tracemalloc.start()
        self.start_time = perf_counter()

    def add(self, label, seconds, count=1):
        """Credit ``seconds`` spread over ``count`` laps to ``label`` (device-side phases are timed in bulk)."""
        if self.record and count > 0:
            self.history.setdefault(label, []).extend([seconds / count] * count)

    def aggregate(self):
        total = 0
        for k, v in self.history.items():
            mean = float(np.mean(np.array(v)))
            total += mean
            print(f'{k}: {mean}')
            if self.memory_usage and k in self.history_mem:
                stored = sum(m[0] for m in self.history_mem[k])
                peak = max(m[1] for m in self.history_mem[k])
                print(f'{k} Memory: Stored {stored} - Peak {peak}')
        print(f'Total: {total}')


# The engine whose jb_pca_project / jb_pca_inverse carry the PCA projection (set by JAMIE while a model is attached to a
# CUDA engine; a module-level hook because preclass objects are pickled inside checkpoints and must stay plain data).
_GPU_PROJECTOR = None


def set_gpu_projector(engine):
    """Route ``preclass`` PCA projections through ``engine`` (None: host numpy / sklearn, e.g. on a machine that only
    inspects checkpoints)."""
    global _GPU_PROJECTOR
    _GPU_PROJECTOR = engine


class preclass:
    """Standardisation (optionally after a fitted PCA) applied at ingest and inverted after ``modal_predict``.

    Same attributes as the reference object so that pickles load either way: ``sample`` (the post-PCA training
    matrix), ``pca`` (anything with transform / inverse_transform), ``axis`` (None: scalar mean/std, used with PCA;
    0: per-feature, used without).  With a linear PCA (``components_`` / ``mean_``, no whitening: sklearn's default,
    the only kind the reference builds, jamie/jamie.py:449-451) and an attached CUDA engine the projection and the
    standardisation run on the GPU (``jb_pca_project`` / ``jb_pca_inverse``: fp32-class split GEMM, <= 2e-6 of the
    float64 host result); anything else (umap, a checkpoint opened without an engine) takes the reference's host path."""

    def __init__(self, sample, pca=None, axis=None):
        self.sample = sample
        self.pca = pca
        self.axis = axis

    def stats(self):
        return self.sample.mean(self.axis), self.sample.std(self.axis)

    def _linear(self):
        p = self.pca
        if p is None or self.axis is not None or getattr(p, 'whiten', False):
            return None
        comp, mean = getattr(p, 'components_', None), getattr(p, 'mean_', None)
        if comp is None or mean is None:
            return None
        return np.asarray(comp), np.asarray(mean)

    def transform(self, X):
        lin = self._linear()
        if lin is not None and _GPU_PROJECTOR is not None and getattr(_GPU_PROJECTOR, 'h', None) and np.ndim(X) == 2:
            m, s = self.stats()
            if np.isfinite(s) and s > 0:
                from .pca_fit import dense_blocks   # scipy.sparse inputs (AnnData.X) are densified 16 k rows at a time
                out = np.empty((X.shape[0], lin[0].shape[0]), np.float64)   # the reference returns float64
                for r0, blk in dense_blocks(X):
                    out[r0:r0 + blk.shape[0]] = _GPU_PROJECTOR.pca_project(blk, lin[0], lin[1], m, s)
                return out
        out = X
        if self.pca is not None:
            out = self.pca.transform(out)
        m, s = self.stats()
        out = out - m
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            with np.errstate(all='ignore'):
                out = out / s
        out[np.isnan(out)] = 0
        return out

    def inverse_transform(self, X):
        lin = self._linear()
        if lin is not None and _GPU_PROJECTOR is not None and getattr(_GPU_PROJECTOR, 'h', None) and np.ndim(X) == 2:
            m, s = self.stats()
            return _GPU_PROJECTOR.pca_inverse(X, lin[0], lin[1], m, s).astype(np.float64)
        m, s = self.stats()
        out = X * s
        out = out + m
        if self.pca is not None:
            out = self.pca.inverse_transform(out)
        return out


class LinearPCA:
    """Minimal PCA (full SVD, sklearn's sign convention) used when scikit-learn is not importable."""

    def __init__(self, n_components):
        self.n_components = n_components

    def fit_transform(self, X):
        X = np.asarray(X, np.float64)
        self.mean_ = X.mean(axis=0)
        U, S, Vt = np.linalg.svd(X - self.mean_, full_matrices=False)
        # sklearn svd_flip (u-based): the largest-magnitude entry of every left singular vector is positive
        idx = np.argmax(np.abs(U), axis=0)
        signs = np.sign(U[idx, range(U.shape[1])])
        signs[signs == 0] = 1
        U *= signs
        Vt *= signs[:, None]
        k = self.n_components
        self.components_ = Vt[:k]
        self.explained_variance_ = (S[:k] ** 2) / max(X.shape[0] - 1, 1)
        return U[:, :k] * S[:k]

    def fit(self, X):
        self.fit_transform(X)
        return self

    def transform(self, X):
        return (np.asarray(X, np.float64) - self.mean_) @ self.components_.T

    def inverse_transform(self, Z):
        return np.asarray(Z, np.float64) @ self.components_ + self.mean_


def make_pca(n_components):
    try:
        from sklearn.decomposition import PCA
        return PCA(n_components=n_components)
    except Exception:  # pragma: no cover
        return LinearPCA(n_components)


# pickled by reference under the reference's module path
identity.__module__ = 'jamie.utilities'
preclass.__module__ = 'jamie.utilities'
time_logger.__module__ = 'jamie.utilities'
