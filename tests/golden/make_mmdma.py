"""End-to-end golden for BASELINE configs[0]: the UNMODIFIED reference (/root/reference/jamie) on its own MMD-MA
simulation files, run in this container (never on the GPU box).

  python tests/golden/make_mmdma.py [--seeds 3] [--threads 4]

Reference recipe: README.md:84-122 -- ``JAMIE(min_epochs=500).fit_transform(dataset=[data1, data2], P=corr)`` on
``examples/data/UnionCom/MMD/s1_mapped{1,2}.txt`` (300 x 2000, 300 x 1000) with the identity prior, then
``modal_predict`` both ways.  BASELINE configs[0] fixes ``pca_dim=None``.  ``use_f_tilde=False`` because estimating F needs
unioncom's geodesic distances, which are not installed here.

Per seed (numpy seed s, torch ``manual_seed`` 666 + s) the script stores the reference's own metrics:
  FOSCTTM          JAMIE.test_closer                      (jamie/jamie.py:892-913)
  LTA (k default)  JAMIE.test_LabelTA                     (jamie/jamie.py:943-961)
  LTA (k = 5)      jamie.evaluation.test_LabelTA          (jamie/evaluation.py:114-132)
  imputation r     mean per-feature Pearson r of modal_predict vs the measured modality, sklearn ``r_regression`` as
                   in ``_plot_correlation``               (jamie/evaluation.py:491-513)
plus epochs run and the final loss_history row.  The seed-to-seed band of these numbers is what the GPU test
(tests/test_gpu_mmdma.py) asserts the B200 implementation falls into.  Output: tests/golden/mmdma.npz (data as fp32).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.ref_harness import import_reference  # noqa: E402

DATA = '/root/reference/examples/data/UnionCom/MMD'


def mean_feature_r(pred, true):
    from sklearn.feature_selection import r_regression
    out = []
    for pr, tr in zip(np.transpose(pred), np.transpose(true)):
        if len(np.unique(tr)) > 1:
            out.append(r_regression(np.reshape(pr, (-1, 1)), tr)[0])
    return float(np.mean(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=3)
    ap.add_argument('--threads', type=int, default=4)
    ap.add_argument('--max-epochs', type=int, default=10000)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    jamie = import_reference()
    import jamie.evaluation as ev
    data1 = np.loadtxt(os.path.join(DATA, 's1_mapped1.txt'))
    data2 = np.loadtxt(os.path.join(DATA, 's1_mapped2.txt'))
    type1 = np.loadtxt(os.path.join(DATA, 's1_type1.txt')).astype(int)
    type2 = np.loadtxt(os.path.join(DATA, 's1_type2.txt')).astype(int)
    corr = np.eye(data1.shape[0], data2.shape[0])
    runs = []
    for s in range(args.seeds):
        np.random.seed(42 + s)
        t0 = time.time()
        jm = jamie.JAMIE(min_epochs=500, pca_dim=None, use_f_tilde=False, manual_seed=666 + s, epoch_DNN=args.max_epochs)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            emb = jm.fit_transform(dataset=[data1.copy(), data2.copy()], P=corr.copy())
            fos = float(jm.test_closer(emb))
            lta, k_def = jm.test_LabelTA(emb, [type1, type2], return_k=True)
            lta5 = float(ev.test_LabelTA(emb, [type1, type2], k=5))
            imp = [np.asarray(jm.modal_predict(data2, 1)), np.asarray(jm.modal_predict(data1, 0))]   # README.md:110-111
        r = [mean_feature_r(imp[0], data1), mean_feature_r(imp[1], data2)]
        run = dict(seed=s, numpy_seed=42 + s, manual_seed=666 + s, epochs=len(jm.loss_history['KL']), foscttm=fos,
                   lta=float(lta), lta_k=int(k_def), lta5=lta5, impute_r=r,
                   final_losses={k: float(v[-1]) for k, v in jm.loss_history.items()}, seconds=time.time() - t0)
        print(json.dumps(run), flush=True)
        runs.append(run)
    np.savez_compressed(os.path.join(HERE, 'mmdma.npz'), data1=data1.astype(np.float32), data2=data2.astype(np.float32),
                        type1=type1, type2=type2, runs=json.dumps(runs),
                        meta=json.dumps(dict(kw=dict(min_epochs=500, pca_dim=None, use_f_tilde=False), source=DATA,
                                             torch=torch.__version__)))


if __name__ == '__main__':
    main()
