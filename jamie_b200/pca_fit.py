"""PCA fit with its O(n d^2) work on the GPU (the reference calls sklearn's ``PCA(n_components=dim)
.fit_transform(data)`` on the host, jamie/jamie.py:436-452, which dominates "Setup" on large inputs).

The algorithm is sklearn's own ``covariance_eigh`` solver (sklearn/decomposition/_pca.py, ``_fit_full``): column means,
the centred Gram matrix C = (X - mean)^T (X - mean) / (n - 1), a symmetric eigendecomposition of the d x d matrix, the
leading eigenvectors as components with the v-based sign convention (``svd_flip(u_based_decision=False)``: the entry of
largest magnitude of every component is positive).  Here the two passes over the [n, d] matrix run on the GPU
(``jb_pca_colsum``, ``jb_pca_gram``: split fp32-class tensor-core GEMM, float64 accumulation over row chunks); only the
d x d eigenproblem (independent of n) is solved on the host with LAPACK.  With the rows sharded over data-parallel ranks
every rank reduces its own rows and the column sums / Gram matrices are all-reduced (one d-vector, one d x d matrix).

The result is returned as a genuine ``sklearn.decomposition.PCA`` object with its fitted attributes set, so that
``preclass`` pickles exactly like the reference's and a checkpoint written here still opens in the reference.
"""
import numpy as np


def _all_reduce_sum(arr, world, group=None):
    """Sum a float64 host array over the `world` ranks of an initialised torch.distributed group (no-op for one rank)."""
    if world <= 1:
        return arr
    try:
        import torch
        import torch.distributed as dist
    except Exception:   # pragma: no cover
        return arr
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend(group) == 'nccl':
        t = t.cuda()
        dist.all_reduce(t, group=group)
        return t.cpu().numpy()
    dist.all_reduce(t, group=group)
    return t.numpy()


BLOCK_ROWS = 16384   # rows densified at a time when the input is a scipy.sparse matrix (AnnData.X)


def is_sparse(X):
    return hasattr(X, 'toarray') and not isinstance(X, np.ndarray)


def dense_blocks(X, lo=0, hi=None, rows=BLOCK_ROWS):
    """float32 row blocks of X[lo:hi]: the whole range at once for an ndarray, BLOCK_ROWS densified rows at a time for a
    scipy.sparse matrix (column sums, Gram matrices and projections are all additive / independent over row blocks, so a
    sparse single-cell matrix never has to exist densely on the host)."""
    hi = X.shape[0] if hi is None else hi
    if not is_sparse(X):
        if hi > lo:
            yield lo, np.ascontiguousarray(X[lo:hi], np.float32)
        return
    Xr = X.tocsr() if hasattr(X, 'tocsr') else X
    for r0 in range(lo, hi, rows):
        r1 = min(hi, r0 + rows)
        yield r0, np.ascontiguousarray(Xr[r0:r1].toarray(), np.float32)


def gram_pca_fit(engine, X, n_components, rank=0, world=1, group=None):
    """Fit on the rows of this rank's contiguous shard of X (ndarray or scipy.sparse), reduce over ranks; returns
    (components [k, d], mean [d], explained_variance [k], total_variance, n_samples), float64."""
    n, d = X.shape
    lo, hi = n * rank // world, n * (rank + 1) // world
    colsum = np.zeros(d)
    for _, blk in dense_blocks(X, lo, hi):
        colsum += engine.pca_colsum(blk)
    colsum = _all_reduce_sum(colsum, world, group)
    mean = colsum / n
    gram = np.zeros((d, d))
    for _, blk in dense_blocks(X, lo, hi):
        gram += engine.pca_gram(blk, mean)
    gram = _all_reduce_sum(gram, world, group)
    gram = (gram + gram.T) * 0.5                      # the tensor-core product is symmetric only to rounding
    C = gram / max(n - 1, 1)
    evals, evecs = np.linalg.eigh(C)                   # ascending
    evals = evals[::-1]
    evecs = evecs[:, ::-1]
    evals = np.maximum(evals, 0.0)                     # sklearn clips the round-off negatives the same way
    total_var = float(evals.sum())
    comps = evecs[:, :n_components].T.copy()
    # svd_flip(u_based_decision=False): sign of the largest-magnitude entry of every component row is positive
    idx = np.argmax(np.abs(comps), axis=1)
    signs = np.sign(comps[np.arange(comps.shape[0]), idx])
    signs[signs == 0] = 1
    comps *= signs[:, None]
    return comps, mean, evals[:n_components].copy(), total_var, n


def fit_sklearn_pca(engine, X, n_components, rank=0, world=1, group=None):
    """``PCA(n_components).fit(X)`` with the GPU Gram route; returns the fitted sklearn object."""
    from sklearn.decomposition import PCA
    comps, mean, ev, total_var, n = gram_pca_fit(engine, X, n_components, rank, world, group)
    d = comps.shape[1]
    pca = PCA(n_components=n_components)
    pca.n_samples_ = n
    pca.n_features_in_ = d
    pca.n_components_ = int(n_components)
    pca.components_ = comps
    pca.mean_ = mean
    pca.explained_variance_ = ev
    pca.explained_variance_ratio_ = ev / total_var if total_var > 0 else np.zeros_like(ev)
    pca.singular_values_ = np.sqrt(ev * max(n - 1, 1))
    rest = min(n, d) - n_components
    pca.noise_variance_ = float((total_var - ev.sum()) / rest) if rest > 0 else 0.0
    pca._fit_svd_solver = 'covariance_eigh'
    return pca


def fit_transform(engine, X, n_components, rank=0, world=1, group=None):
    """(fitted sklearn PCA, sample = the projected training matrix [n, k] float64): what ``pca.fit_transform(data)`` gives
    the reference (jamie/jamie.py:451); the projection is ``jb_pca_project`` without standardisation."""
    pca = fit_sklearn_pca(engine, X, n_components, rank, world, group)
    sample = np.empty((X.shape[0], int(n_components)), np.float64)
    for r0, blk in dense_blocks(X):
        sample[r0:r0 + blk.shape[0]] = engine.pca_project(blk, pca.components_, pca.mean_, 0.0, 1.0)
    return pca, sample
