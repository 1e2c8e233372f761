"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/jamie) in this container.

Run:  python tests/golden/make_golden.py          (needs /root/reference; never run on the GPU box)

Each fixture holds, for a small configuration: the inputs (datasets, P, match_result, constructor kwargs), the initial
parameters, and for every optimizer step of the reference's own ``project_jamie`` loop the recorded randomness
(numpy draws, dropout masks, eps) and results (forward tensors, pre-clip gradients, total norm, post-step parameters
and BatchNorm buffers), plus the end-of-run outputs (``fit_transform`` embeddings, ``loss_history``, ``modal_predict``,
``transform_one``).  ``tests/test_oracle_golden.py`` replays them through ``oracle/jamie_oracle.py``.
"""
import io
import json
import os
import sys
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.ref_harness import Tap, import_reference  # noqa: E402


def synth(n, d, seed, latent=4):
    rng = np.random.default_rng(seed)
    t = rng.normal(size=(n, latent))
    out = []
    for di in d:
        A = rng.normal(size=(latent, latent))
        Bm = rng.normal(size=(latent, di))
        out.append((np.tanh(t @ A) @ Bm + 0.1 * rng.normal(size=(n, di))).astype(np.float64))
    return out


CASES = {
    # name: (n0, n1, dims, kwargs, P-kind, F-kind)
    'diag_drop': dict(n=(40, 40), d=(72, 40), P='none', F='zero',
                      kw=dict(output_dim=8, batch_size=32, pca_dim=None, epoch_DNN=3, min_epochs=2, use_f_tilde=False)),
    'rep_F': dict(n=(48, 48), d=(40, 12), P='eye', F='dense',
                  kw=dict(output_dim=6, batch_size=48, pca_dim=None, epoch_DNN=3, min_epochs=4, dropout=0.4,
                          PF_Ratio=0.7, loss_weights=[1, 2, 0.5, 3])),
    'hybrid': dict(n=(50, 50), d=(30, 26), P='half', F='zero',
                   kw=dict(output_dim=8, batch_size=32, pca_dim=None, epoch_DNN=3, min_epochs=0, dropout=0.3,
                           use_f_tilde=False)),
    'zeros_unequal': dict(n=(44, 36), d=(20, 24), P='none', F='zero',
                          kw=dict(output_dim=4, batch_size=32, pca_dim=None, epoch_DNN=2, min_epochs=2,
                                  use_f_tilde=False)),
    'pca': dict(n=(60, 60), d=(100, 80), P='none', F='zero',
                kw=dict(output_dim=8, batch_size=32, pca_dim=[16, 12], epoch_DNN=3, min_epochs=2, dropout=0.5,
                        use_f_tilde=False)),
    # dist_method='cosine' (the other live branch of sim_diff_func, jamie/jamie.py:485-494); a partially matched prior and
    # F so that c != z and every latent term is exercised
    'cosine': dict(n=(56, 56), d=(36, 28), P='half', F='dense',
                   kw=dict(output_dim=8, batch_size=32, pca_dim=None, epoch_DNN=3, min_epochs=1, dropout=0.3,
                           dist_method='cosine', PF_Ratio=0.6, loss_weights=[1, 1, 4, 2])),
    'multibatch': dict(n=(100, 100), d=(24, 16), P='none', F='zero',
                       kw=dict(output_dim=8, batch_size=32, pca_dim=None, epoch_DNN=2, min_epochs=2, dropout=0.2,
                               use_f_tilde=False)),
}


def build_case(name, spec):
    jamie = import_reference()
    n0, n1 = spec['n']
    seed = abs(hash(name)) % 1000 if False else sum(ord(c) for c in name)
    if n0 == n1:
        data = synth(n0, spec['d'], seed)
    else:
        data = [synth(n0, spec['d'][:1], seed)[0], synth(n1, spec['d'][1:], seed + 1)[0]]
    rng = np.random.default_rng(seed + 7)
    P = None
    if spec['P'] == 'eye':
        P = np.eye(n0)
    elif spec['P'] == 'half':
        P = np.diag((rng.random(n0) < 0.5).astype(np.float64))
    kw = dict(spec['kw'])
    match_result = None
    if spec['F'] == 'dense':
        Fm = rng.random((n0, n1)) * (rng.random((n0, n1)) < 0.3)
        match_result = [Fm]
        kw['match_result'] = match_result
    tap = Tap()
    out = {}
    with tap:
        np.random.seed(42)
        jm = jamie.JAMIE(model_class=tap.model_class(), **kw)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            emb = jm.fit_transform(dataset=[d.copy() for d in data], P=None if P is None else P.copy())
    # end-of-run outputs (eval mode was set by project_jamie, jamie.py:794)
    pred = [jm.modal_predict(data[i], i) for i in range(2)]
    t_one = [jm.transform_one(data[i], i) for i in range(2)]
    pre = [jm.model.preprocessing[i](data[i]) for i in range(2)]
    out['meta'] = json.dumps(dict(name=name, n=[n0, n1], d=list(spec['d']), kw=kw if match_result is None else
                                  {k: v for k, v in kw.items() if k != 'match_result'},
                                  P=spec['P'], F=spec['F'], sampling_method=jm.sampling_method,
                                  batch_size=int(jm.batch_size), col=[int(c) for c in jm.col],
                                  dropout=float(tap.model.encoders[0][3].p),
                                  param_names=tap.param_names, n_steps=len(tap.steps), stdout=buf.getvalue()[-2000:]))
    for i in range(2):
        out[f'data{i}'] = data[i]
        out[f'emb{i}'] = emb[i]
        out[f'pred{i}'] = np.asarray(pred[i])
        out[f'tone{i}'] = t_one[i]
        out[f'pre{i}'] = np.asarray(pre[i])
    if P is not None:
        out['P'] = P
    if match_result is not None:
        out['F'] = match_result[0]
    if kw.get('pca_dim') is not None:
        for i in range(2):
            pca = jm.model.preprocessing[i].__self__.pca
            out[f'pca_components{i}'] = pca.components_
            out[f'pca_mean{i}'] = pca.mean_
    for k, p in enumerate(tap.init_params):
        out[f'init/p{k}'] = p
    for n, b in tap.init_buffers.items():
        out[f'init/b/{n}'] = b
    for s, st in enumerate(tap.steps):
        for k, a in enumerate(st['choice']):
            out[f's{s}/choice{k}'] = a
        for k, a in enumerate(st['rand']):
            out[f's{s}/rand{k}'] = a
        for k, a in enumerate(st['masks']):
            out[f's{s}/mask{k}'] = np.packbits(a.astype(bool), axis=None)
            out[f's{s}/maskshape{k}'] = np.array(a.shape)
        for k, a in enumerate(st['eps']):
            out[f's{s}/eps{k}'] = a
        mo = st['model_out']
        for key in ('x', 'z', 'c', 'xhat', 'mu'):
            for i in range(2):
                out[f's{s}/{key}{i}'] = mo[key][i]
        out[f's{s}/corr'] = mo['corr']
        out[f's{s}/logvars'] = mo['logvars']
        for k, g in enumerate(st['grads']):
            out[f's{s}/g{k}'] = g
        out[f's{s}/total_norm'] = np.array(st['total_norm'])
        for k, p in enumerate(st['params_after']):
            out[f's{s}/p{k}'] = p
        for n, b in st['buffers_after'].items():
            out[f's{s}/b/{n}'] = b
    for k, v in jm.loss_history.items():
        out[f'loss_history/{k}'] = np.array(v, dtype=np.float64)
    return out


def build_checkpoint():
    """A checkpoint written by the reference's own save_model (plain reference classes, PCA + standardisation inside)
    with the predictions the reference makes from it: the cross-loading fixture."""
    jamie = import_reference()
    data = synth(90, (60, 45), 77)
    np.random.seed(42)
    jm = jamie.JAMIE(output_dim=8, batch_size=32, pca_dim=[20, 14], epoch_DNN=4, min_epochs=2, dropout=0.3,
                     use_f_tilde=False)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        emb = jm.fit_transform(dataset=[d.copy() for d in data])
    path = os.path.join(HERE, 'ref_checkpoint.h5')
    jm.save_model(path)
    out = {f'data{i}': data[i] for i in range(2)}
    for i in range(2):
        out[f'emb{i}'] = emb[i]
        out[f'pred{i}'] = np.asarray(jm.modal_predict(data[i], i))
        out[f'tone{i}'] = jm.transform_one(data[i], i)
    np.savez_compressed(os.path.join(HERE, 'ref_checkpoint_io.npz'), **out)
    print('checkpoint', os.path.getsize(path) // 1024, 'KiB')


def main():
    torch.set_num_threads(4)
    only = sys.argv[1:]          # `python make_golden.py cosine` (re)generates the named cases only
    if not only:
        build_checkpoint()
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        out = build_case(name, spec)
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **out)
        print(name, 'steps', json.loads(out['meta'])['n_steps'], os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
