"""Distances and the correspondence estimate F: what ``fit_transform`` runs before training when
``use_f_tilde=True`` and no ``match_result`` is supplied (jamie/jamie.py:160-175).

* ``distance_function`` / ``compute_distances`` -- jamie/jamie.py:839-890: per-dataset n x n distance matrix in the chosen
  ``distance_mode`` (host: sklearn / scipy, exactly the calls the reference makes).  The default mode 'geodesic' calls the
  third-party ``unioncom.utils.geodesic_distances`` (unioncom==0.4.0, NOT vendored in the reference): restated here from
  the published package -- kNN graph (k from 5 upwards in steps of 2 until connected or k > max(kmax, 0.01 n)), all-pairs
  shortest paths, unreachable pairs set to twice the largest finite distance.  PARITY UNPINNED for that function (no
  reference test or fixture covers it); the other modes and Prime_Dual are pinned by tests/golden/prime_dual.npz.
* ``prime_dual`` -- jamie/jamie.py:314-414: UnionCom's primal-dual Adam iteration for F on the GPU.  Dense n x n fp32
  algebra: four GEMMs per iteration through torch / cuBLAS (plain library GEMMs), everything else element-wise; the
  rank-one products of the reference (``Mu 1^T``, ``1 Lambda^T``, ``F 1 1^T``, ``1 1^T F``) are written as broadcasts of
  row / column sums.  Same update order, constants and printed lines as the reference.
"""
import warnings

import numpy as np


def geodesic_distances(X, kmax):
    """unioncom.utils.geodesic_distances (unioncom==0.4.0), restated; see the module docstring (parity unpinned)."""
    import scipy.sparse.csgraph as csg
    from sklearn.neighbors import NearestNeighbors
    X = np.asarray(X)
    n = len(X)
    k = 5

    def graph(k_):
        nb = NearestNeighbors(n_neighbors=min(k_, n), metric='euclidean').fit(X)
        return nb.kneighbors_graph(X, mode='distance')

    knn = graph(k)
    while csg.connected_components(knn, directed=False)[0] != 1:
        if k > max(kmax, 0.01 * n):
            break
        k += 2
        knn = graph(k)
    dist = csg.shortest_path(knn, method='D', directed=False)
    finite_max = np.nanmax(dist[dist != np.inf])
    dist[dist > finite_max] = 2 * finite_max
    return dist


def distance_function(distance_mode, kmax):
    """The per-dataset distance callable of jamie/jamie.py:851-884."""
    if distance_mode == 'geodesic':
        def fn(df):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                return np.array(geodesic_distances(df, kmax))
    elif distance_mode == 'spearman':
        def fn(df):
            from scipy import stats
            if df.shape[0] == 1:
                d = np.array([0])
            else:
                d, _ = stats.spearmanr(df, axis=1)
                if np.isnan(d).any():
                    raise Exception('Data is not well conditioned for spearman method '
                                    '(scipy.stats.spearmanr returned ``np.nan``)')
            if len(np.shape(d)) == 0:
                d = np.array([[1, d], [d, 1]])
            return (1 - np.array(d)) / 2
    elif distance_mode == 'pearson':
        def fn(df):
            if df.shape[0] == 1:
                return np.array([0])
            d = np.corrcoef(df.toarray() if hasattr(df, 'toarray') else df)
            if len(np.shape(d)) == 0:
                d = np.array([[1, d], [d, 1]])
            return (1 - np.array(d)) / 2
    else:
        def fn(df):
            from sklearn.metrics import pairwise_distances
            return pairwise_distances(df, metric=distance_mode)
    return fn


def prime_dual(Kx, Ky, dx, dy, epoch_pd=2000, epsilon=0.01, rho=10, delay=0, log_pd=100, verbose=True, device='cuda'):
    """F (m x n, fp32 ndarray) minimising ||a Kx - F Ky F^T|| under soft row / column constraints: jamie/jamie.py:314-414."""
    import torch
    if np.shape(Kx) == (1, 1) and np.shape(Ky) == (1, 1):
        warnings.warn('1x1 distance matrix, escaping...')
        return np.ones((1, 1), np.float32)
    N = int(max(np.shape(Kx)[0], np.shape(Ky)[0]))
    Kx = torch.from_numpy(np.asarray(Kx) / N).float().to(device)
    Ky = torch.from_numpy(np.asarray(Ky) / N).float().to(device)
    a = float(np.sqrt(dy / dx))
    m, n = Kx.shape[0], Ky.shape[0]
    F = torch.zeros((m, n), device=device)
    lam = torch.zeros((1, n), device=device)    # Lambda^T
    mu = torch.zeros((m, 1), device=device)
    S = torch.zeros((1, n), device=device)      # S^T
    m1 = torch.zeros_like(F)
    m2 = torch.zeros_like(F)
    b1, b2, delta = 0.9, 0.999, 10e-8
    trace_kk = torch.sum(Kx * Kx.t())           # trace(Kx Kx)
    for i in range(1, epoch_pd + 1):
        FKy = F @ Ky
        grad = (4.0 * (FKy @ (F.t() @ FKy)) - (4.0 * a) * (Kx @ FKy) + mu + lam
                + rho * (F.sum(dim=1, keepdim=True) + (F.sum(dim=0, keepdim=True) + (S - 2.0))))
        m1 = b1 * m1 + (1 - b1) * grad
        m2 = b2 * m2 + (1 - b2) * grad * grad
        step = (m1 / (1 - b1 ** i)) / (torch.sqrt(m2 / (1 - b2 ** i)) + delta)
        F = (1 - epsilon) * F + epsilon * torch.clamp(F - step, min=0)
        col = F.sum(dim=0, keepdim=True)        # (F^T 1)^T
        S = (1 - epsilon) * S + epsilon * torch.clamp(S - (lam + rho * (col - 1.0 + S)), min=0)
        mu = mu + epsilon * (F.sum(dim=1, keepdim=True) - 1.0)
        lam = lam + epsilon * (col - 1.0 + S)
        if i >= delay:
            FKyFt = (F @ Ky) @ F.t()
            a = torch.sum(Kx * FKyFt.t()) / trace_kk          # trace(Kx F Ky F^T) / trace(Kx Kx)
        if verbose and i % log_pd == 0:
            err = torch.norm(a * Kx - (F @ Ky) @ F.t())
            print('epoch:[{:d}/{:d}] err:{:.4f} alpha:{:.4f}'.format(i, epoch_pd, float(err), float(a)))
    return F.detach().cpu().numpy()
