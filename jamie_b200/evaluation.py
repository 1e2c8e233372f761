"""The three acceptance metrics the reference's notebooks report, used here as end-to-end parity metrics (small n, CPU).

Reference definitions: FOSCTTM jamie/evaluation.py:65-85 (and jamie/jamie.py:892-913), label-transfer accuracy
jamie/evaluation.py:114-132 (class method: jamie/jamie.py:943-961), per-feature imputation correlation
jamie/evaluation.py:491-513.
"""
import numpy as np


def foscttm(a, b):
    """Fraction of samples closer than the true match, averaged over both directions (lower is better)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    n = d.shape[0]
    true = np.diag(d)
    frac0 = ((d < true[:, None]).sum(axis=1) / (n - 1)).mean()
    frac1 = ((d < true[None, :]).sum(axis=0) / (n - 1)).mean()
    return float((frac0 + frac1) / 2)


def label_transfer_accuracy(integrated_data, datatype, k=None, return_k=False):
    """kNN classifier fitted on modality 1's embedding and labels, scored on modality 0's."""
    emb0, emb1 = [np.asarray(e, np.float64) for e in integrated_data[:2]]
    y0, y1 = [np.asarray(t).ravel() for t in datatype[:2]]
    if k is None:
        # class method default: 20% of the average class size (jamie/jamie.py:949-953)
        _, counts = np.unique(y1, return_counts=True)
        k = max(1, int(.2 * counts.mean()))
    d = ((emb0[:, None, :] - emb1[None, :, :]) ** 2).sum(-1)
    nn = np.argsort(d, axis=1, kind='stable')[:, :k]
    labels = y1[nn]
    classes = np.unique(y1)
    votes = np.stack([(labels == c).sum(axis=1) for c in classes], axis=1)
    pred = classes[np.argmax(votes, axis=1)]
    acc = float((pred == y0).mean())
    return (acc, k) if return_k else acc


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
