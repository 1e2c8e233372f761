"""Data-parallel host logic on CPU: two gloo ranks shard the cells, agree on the epoch length and reduce a flat
gradient buffer + loss tail exactly as the GPU path does with NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from jamie_b200.jamie import PriorSpec, _dist_info
    assert _dist_info() == (rank, world)
    rows = [1001, 1001]
    lo = [(r * rank) // world for r in rows]
    hi = [(r * (rank + 1)) // world for r in rows]
    m = (np.arange(rows[0]) % 2 == 0).astype(np.float32)
    prior = PriorSpec(m[lo[0]:hi[0]], [hi[0] - lo[0], hi[1] - lo[1]])
    # the flat-buffer reduction: sum then divide by world inside the update
    g = torch.full((16,), float(rank + 1))
    dist.all_reduce(g)
    len_dataloader = int(max(rows) / (128 * world))
    # row-sharded inference through the API helper (JAMIE.modal_predict / transform_one / the final encode): every rank
    # passes the same rows, computes only its own range and receives the full result
    from jamie_b200.jamie import shard_rows_apply
    full = np.arange(7 * 3, dtype=np.float32).reshape(7, 3)      # 7 rows over 2 ranks: ranges of 3 and 4 rows
    seen = []
    got = shard_rows_apply(lambda x: (seen.append(len(x)), x * 2 + rank * 0)[1], full)
    assert seen == [3 if rank == 0 else 4] and np.array_equal(got, full * 2)
    # PCA fit with the rows sharded over the ranks (jamie_b200/pca_fit.py): each rank reduces its own rows (here a numpy
    # stand-in for jb_pca_colsum / jb_pca_gram), column sums and Gram matrices are all-reduced, every rank solves the same
    # d x d eigenproblem -> the same components as one rank over all rows, and as sklearn's own fit
    from jamie_b200 import pca_fit

    class HostPasses:
        def pca_colsum(self, X):
            return np.asarray(X, np.float64).sum(0)

        def pca_gram(self, X, mean):
            Xc = np.asarray(X, np.float64) - mean
            return Xc.T @ Xc

    rng = np.random.default_rng(3)
    X = (rng.normal(size=(301, 6)) * np.linspace(3, 0.5, 6)) @ rng.normal(size=(6, 20)) + rng.normal(size=20)
    X = X.astype(np.float32)
    comps, mean, ev, tot, n = pca_fit.gram_pca_fit(HostPasses(), X, 4, rank, world)
    comps1, mean1, ev1, tot1, n1 = pca_fit.gram_pca_fit(HostPasses(), X, 4, 0, 1)
    from sklearn.decomposition import PCA
    ref = PCA(n_components=4, svd_solver='full').fit(X.astype(np.float64))
    pca_ok = (n == n1 == 301 and np.allclose(comps, comps1, atol=1e-10) and np.allclose(ev, ev1, rtol=1e-10)
              and np.allclose(comps, ref.components_, atol=1e-8) and np.allclose(ev, ref.explained_variance_, rtol=1e-8)
              and np.allclose(mean, ref.mean_, atol=1e-10))
    q.put((rank, lo, hi, prior.sampling_method, prior.corr_samples.tolist(), g.tolist(), len_dataloader, bool(pca_ok)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduce():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, m0, c0, g0, l0, p0), (r1, lo1, hi1, m1, c1, g1, l1, p1) = res
    assert p0 and p1                                                          # sharded PCA fit == single-rank fit == sklearn
    assert lo0 == [0, 0] and hi0 == lo1 and hi1 == [1001, 1001]          # contiguous, disjoint, covering
    assert m0 == m1 == 'hybrid'
    assert c0 == [[0, 0], [2, 2]] and c1 == [[0, 0], [2, 2]]                 # shard-local indices (500 is even)
    assert g0 == g1 == [3.0] * 16
    assert l0 == l1 == 3
