"""Throughput of the rows SURVEY.md 8 marks "next" (N3 PCA fit, N4 evaluation metrics) on the GPU, with the host
implementation the reference calls timed beside them on a bounded sample. Prints one markdown table (profiles/aux_r2.md).
Timings are wall clock around the C-ABI calls, HOST arrays in and out (upload, kernels, download)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from jamie_b200 import evaluation as E, pca_fit
from jamie_b200.engine import Engine
from oracle import metrics_oracle as MO

torch.cuda.set_device(0)
rows = []


def timed(f, reps=3):
    f()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = f(); t.append(time.perf_counter() - t0)
    return min(t), out


rng = np.random.default_rng(0)
# ---- FOSCTTM / label transfer on embeddings of n cells, L = 32
for n, n_host in ((20000, 20000), (100000, None)):
    a = rng.normal(size=(n, 32)).astype(np.float32)
    b = (a + 1.5 * rng.normal(size=a.shape)).astype(np.float32)
    y = rng.integers(0, 12, size=n)
    tg, fos = timed(lambda: E.test_closer([a, b], verbose=False))
    th = None
    if n_host:
        t0 = time.perf_counter(); fos_h = MO.test_closer([a, b], verbose=False); th = time.perf_counter() - t0
        assert fos == fos_h, (fos, fos_h)
    rows.append((f'FOSCTTM, n = {n}, L = 32 (2 n^2 distances)', tg, th, f'{2 * n * n / tg / 1e9:.1f} G pairs/s'))
    tg, acc = timed(lambda: E.test_LabelTA([a, b], [y, y], k=None, verbose=False), reps=2)
    th = None
    if n_host:
        t0 = time.perf_counter(); acc_h = MO.test_LabelTA([a, b], [y, y], k=None, verbose=False); th = time.perf_counter() - t0
        assert acc == acc_h, (acc, acc_h)
    rows.append((f'label transfer, n = {n}, k = {E.default_k([y, y])} (n^2 distances + selection)', tg, th, f'{n * n / tg / 1e9:.1f} G pairs/s'))
# ---- per-feature Pearson r of an imputed matrix
x = rng.normal(size=(100000, 512)).astype(np.float32); yv = (0.5 * x + rng.normal(size=x.shape)).astype(np.float32)
tg, r = timed(lambda: E.imputation_correlation(x, yv))
t0 = time.perf_counter(); r_h = MO.imputation_correlation(x, yv); th = time.perf_counter() - t0
assert np.abs(r - r_h).max() < 1e-9
rows.append(('per-feature Pearson r, 100000 x 512', tg, th, f'{2 * x.nbytes / tg / 1e9:.1f} GB/s of host input'))
del x, yv
# ---- PCA fit (the reference: sklearn PCA(n_components).fit_transform on the host)
from sklearn.decomposition import PCA
eng = Engine([8, 8], 2, 8, 0.0)
for n, d, k in ((100000, 1000, 128), (200000, 2000, 512)):
    lat = rng.normal(size=(n, 64)).astype(np.float32) * np.linspace(3, 0.3, 64, dtype=np.float32)
    X = lat @ rng.normal(size=(64, d)).astype(np.float32) + 0.1 * rng.normal(size=(n, d)).astype(np.float32)
    tg, (pca, sample) = timed(lambda: pca_fit.fit_transform(eng, X, k), reps=1)
    th = None
    if n <= 100000:
        t0 = time.perf_counter(); ref = PCA(n_components=k).fit(X); th = time.perf_counter() - t0
        cos = np.abs(np.sum(pca.components_[:32] * ref.components_[:32], axis=1)).min()
        note = f'min |cos| of the first 32 components vs sklearn ({ref._fit_svd_solver}) {cos:.6f}'
    else:
        note = 'sklearn not timed at this size'
    rows.append((f'PCA fit + projection, {n} x {d} -> {k}', tg, th, f'{2.0 * n * d * d / tg / 1e12:.1f} TFLOP/s of Gram work; {note}'))
    del X, lat
eng.close()
print('| task | GPU path (s, host arrays in / out) | host implementation (s) | note |')
print('|---|---|---|---|')
for name, tg, th, note in rows:
    print(f'| {name} | {tg:.3f} | {"-" if th is None else f"{th:.2f}"} | {note} |')
