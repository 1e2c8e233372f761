"""Generates tests/golden/metrics.npz by running the UNMODIFIED reference's evaluation functions
(/root/reference/jamie/evaluation.py: test_closer :65-85, test_LabelTA :114-132; the class methods JAMIE.test_closer /
JAMIE.test_LabelTA, jamie/jamie.py:892-961; the per-feature correlation of _plot_correlation, sklearn r_regression,
:491-513) on small seeded embeddings.  Run:  python tests/golden/make_metrics.py   (needs /root/reference).
``tests/test_oracle_golden.py`` pins oracle/metrics_oracle.py with it, ``tests/test_gpu_metrics.py`` the GPU metrics."""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.ref_harness import import_reference  # noqa: E402


def main():
    jamie = import_reference()
    import jamie.evaluation as ev
    from sklearn.feature_selection import r_regression
    out, meta = {}, []
    cases = [dict(n0=120, n1=120, L=8, ncls=3, noise=3.5, seed=1), dict(n0=200, n1=150, L=16, ncls=5, noise=5.0, seed=2),
             dict(n0=90, n1=90, L=4, ncls=2, noise=2.5, seed=3)]
    for c, spec in enumerate(cases):
        rng = np.random.default_rng(spec['seed'])
        y0 = rng.integers(0, spec['ncls'], size=spec['n0'])
        y1 = rng.integers(0, spec['ncls'], size=spec['n1']) if spec['n0'] != spec['n1'] else y0.copy()
        centers = rng.normal(size=(spec['ncls'], spec['L'])) * 2
        e0 = (centers[y0] + spec['noise'] * rng.normal(size=(spec['n0'], spec['L']))).astype(np.float32)
        e1 = ((e0 if spec['n0'] == spec['n1'] else centers[y1]) + spec['noise'] * rng.normal(size=(spec['n1'], spec['L']))).astype(np.float32)
        rec = dict(spec)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            jm = jamie.JAMIE()
            if spec['n0'] == spec['n1']:
                rec['foscttm'] = float(ev.test_closer([e0, e1]))
                rec['foscttm_method'] = float(jm.test_closer([e0, e1]))
            for k in (1, 5, 17):
                rec[f'lta_k{k}'] = float(ev.test_LabelTA([e0, e1], [y0, y1], k=k))
            acc, kdef = jm.test_LabelTA([e0, e1], [y0, y1], return_k=True)
            rec['lta_default'], rec['k_default'] = float(acc), int(kdef)
        rec['stdout'] = buf.getvalue()[-400:]
        # per-feature correlation as _plot_correlation computes it (jamie/evaluation.py:491-513)
        x = rng.normal(size=(spec['n0'], 12)).astype(np.float32) * 2 + 3
        yv = (0.7 * x + rng.normal(size=x.shape)).astype(np.float32)
        yv[:, 5] = 2.0      # a constant feature: skipped by the reference (len(np.unique(tr)) > 1)
        r = [float(r_regression(np.reshape(pr, (-1, 1)), tr)[0]) if len(np.unique(tr)) > 1 else float('nan')
             for pr, tr in zip(np.transpose(x), np.transpose(yv))]
        out[f'c{c}/e0'], out[f'c{c}/e1'], out[f'c{c}/y0'], out[f'c{c}/y1'] = e0, e1, y0, y1
        out[f'c{c}/x'], out[f'c{c}/y'], out[f'c{c}/r'] = x, yv, np.array(r)
        meta.append(rec)
        print(json.dumps({k: v for k, v in rec.items() if k != 'stdout'}))
    out['meta'] = json.dumps(meta)
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), **out)
    print('metrics.npz', os.path.getsize(os.path.join(HERE, 'metrics.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
