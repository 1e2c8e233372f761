"""TEST INFRASTRUCTURE ONLY -- never imported by the product (jamie_b200/) or by anything that runs on the GPU box.

Imports the *unmodified* reference (``/root/reference/jamie``) in this container so that golden fixtures can be
generated from it (``tests/golden/make_golden.py``) and the numpy oracle (``oracle/jamie_oracle.py``) can be pinned.

Two parts:

1. ``install_stubs()``: ``sys.modules`` stubs for third-party packages that are not installed here and that the
   reference imports at module level (matplotlib, seaborn, adjustText, brokenaxes, umap, anndata, unioncom).
   ``unioncom==0.4.0`` is not vendored under /root/reference: its ``UnionCom.__init__`` defaults and
   ``init_random_seed`` are restated from the published package (PARITY UNPINNED for those defaults: no reference
   test covers them; the values visible in the reference's stored notebook logs -- epoch_pd=2000, manual_seed=666 --
   agree).  The stubs contain no reference code.
2. ``Tap``: records, per optimizer step of the real ``JAMIE.project_jamie`` loop (jamie/jamie.py:546-749), every
   source of randomness and every result needed to replay the step: ``np.random.choice`` / ``np.random.rand`` draws
   (jamie.py:556-578), the 8 dropout masks (model.py:154,164,195,200), the reparameterisation eps (model.py:239-240),
   the pre-clip gradients and total norm (jamie.py:739), the post-step parameters and BatchNorm buffers (jamie.py:740).
"""
import random
import sys
import types

import numpy as np
import torch

import os

# The reference package is imported from /root/reference when that exists (the build container) and otherwise from
# oracle/_ref (a byte-for-byte copy of /root/reference/jamie made by __graft_entry__.build(); git-ignored, it only
# travels to the GPU box so that bench.py --impl reference can time the real reference there).
_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REFERENCE_ROOT = '/root/reference' if os.path.isdir('/root/reference/jamie') else _REF_COPY


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'jamie'))


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _ga(n):
    if n.startswith('__'):
        raise AttributeError(n)
    return _Any()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class UnionComStub(object):
    """Attribute defaults of unioncom 0.4.0 ``UnionCom.__init__`` (restated; see module docstring)."""

    def __init__(self, integration_type='MultiOmics', epoch_pd=2000, epoch_DNN=100, epsilon=0.01, lr=0.001,
                 batch_size=100, rho=10, beta=1, perplexity=30, log_DNN=10, log_pd=100, manual_seed=666, delay=0,
                 kmax=40, output_dim=32, distance_mode='geodesic', project_mode='tsne'):
        self.integration_type = integration_type
        self.epoch_pd = epoch_pd
        self.epoch_DNN = epoch_DNN
        self.epsilon = epsilon
        self.lr = lr
        self.batch_size = batch_size
        self.rho = rho
        self.beta = beta
        self.perplexity = perplexity
        self.log_DNN = log_DNN
        self.log_pd = log_pd
        self.manual_seed = manual_seed
        self.delay = delay
        self.kmax = kmax
        self.output_dim = output_dim
        self.distance_mode = distance_mode
        self.project_mode = project_mode


def _init_random_seed(manual_seed):
    seed = random.randint(1, 10000) if manual_seed is None else manual_seed
    print("use random seed: {}".format(seed))
    random.seed(seed)
    torch.manual_seed(seed)


def _unavailable(*a, **k):
    raise NotImplementedError('unioncom is not installed; run with use_f_tilde=False or a supplied match_result')


def install_stubs():
    if 'jamie' in sys.modules and getattr(sys.modules['jamie'], '__file__', '').startswith(REFERENCE_ROOT):
        return
    mpl = _mod('matplotlib')
    _mod('matplotlib.pyplot', __getattr__=_ga)
    _mod('matplotlib.collections', PatchCollection=_Any)
    mpl.pyplot = sys.modules['matplotlib.pyplot']
    mpl.collections = sys.modules['matplotlib.collections']
    _mod('seaborn', __getattr__=_ga)
    _mod('adjustText', adjust_text=lambda *a, **k: None)
    _mod('brokenaxes', brokenaxes=_Any)
    _mod('umap', UMAP=_Any)

    class AnnData:
        pass

    ad = _mod('anndata')
    core = _mod('anndata._core')
    adm = _mod('anndata._core.anndata', AnnData=AnnData)
    ad._core = core
    core.anndata = adm
    uc = _mod('unioncom')
    ucm = _mod('unioncom.UnionCom', UnionCom=UnionComStub)
    uc.UnionCom = ucm
    _mod('unioncom.utils', geodesic_distances=_unavailable, init_random_seed=_init_random_seed,
         joint_probabilities=_unavailable)
    for name in [k for k in sys.modules if k == 'jamie' or k.startswith('jamie.')]:
        del sys.modules[name]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
    """Returns the reference's ``jamie`` package (imported from /root/reference, or from the oracle/_ref copy)."""
    install_stubs()
    import jamie  # noqa
    assert jamie.__file__.startswith(REFERENCE_ROOT), jamie.__file__
    return jamie


class Tap:
    """Context manager that records one reference training run step by step."""

    def __init__(self, max_steps=None):
        self.steps = []       # list of dict per optimizer step
        self.cur = None
        self.init_params = None
        self.param_names = None
        self.model = None
        self.max_steps = max_steps

    # ---- helpers
    def _new_step(self):
        self.cur = {'choice': [], 'rand': [], 'masks': [], 'eps': [], 'model_out': None}

    def model_class(self):
        """A ``model_class`` for the reference constructor: the reference's own edModelVar with recording Dropout."""
        jamie = import_reference()
        tap = self

        class RecDropout(torch.nn.Module):
            def __init__(self, p):
                super().__init__()
                self.p = p

            def forward(self, x):
                if not self.training:
                    return x
                if self.p == 0:
                    mask = torch.ones_like(x)
                    tap.cur['masks'].append(mask.numpy().astype(np.uint8))
                    return x
                mask = torch.bernoulli(torch.full_like(x, 1 - self.p))
                tap.cur['masks'].append(mask.numpy().astype(np.uint8))
                return x * mask / (1 - self.p)

        class TappedModel(jamie.model.edModelVar):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                for seq in list(self.encoders) + list(self.decoders):
                    for idx, child in list(seq.named_children()):
                        if isinstance(child, torch.nn.Dropout):
                            seq[int(idx)] = RecDropout(child.p)
                tap.model = self

            def forward(self, *X, corr):
                out = super().forward(*X, corr=corr)
                if self.training and tap.cur is not None:
                    zs, combined, xhat, mus, logvars = out
                    tap.cur['model_out'] = {
                        'x': [x.detach().numpy().copy() for x in X],
                        'corr': corr.detach().numpy().copy(),
                        'z': [t.detach().numpy().copy() for t in zs],
                        'c': [t.detach().numpy().copy() for t in combined],
                        'xhat': [t.detach().numpy().copy() for t in xhat],
                        'mu': [t.detach().numpy().copy() for t in mus],
                        'logvars': logvars.detach().numpy().copy(),
                    }
                return out

        # pickling (save_model) needs an importable class
        TappedModel.__module__ = __name__
        TappedModel.__qualname__ = 'TappedModel'
        globals()['TappedModel'] = TappedModel
        return TappedModel

    def __enter__(self):
        jamie = import_reference()
        tap = self
        jm = jamie.jamie
        self._orig = {
            'choice': np.random.choice, 'rand': np.random.rand,
            'std_normal': torch.distributions.normal._standard_normal,
            'clip': torch.nn.utils.clip_grad_norm_, 'Adam': jm.optim.Adam,
        }
        self._new_step()

        def choice(*a, **k):
            r = tap._orig['choice'](*a, **k)
            tap.cur['choice'].append(np.array(r).copy())
            return r

        def rand(*a, **k):
            r = tap._orig['rand'](*a, **k)
            tap.cur['rand'].append(np.array(r).copy())
            return r

        def std_normal(shape, dtype, device):
            e = tap._orig['std_normal'](shape, dtype, device)
            tap.cur['eps'].append(e.numpy().copy())
            return e

        def clip(parameters, max_norm, *a, **k):
            params = list(parameters)
            tap.cur['grads'] = [p.grad.detach().numpy().copy() for p in params]
            tn = tap._orig['clip'](params, max_norm, *a, **k)
            tap.cur['total_norm'] = float(tn)
            return tn

        class RecAdam(self._orig['Adam']):
            def __init__(self, params, *a, **k):
                params = list(params)
                super().__init__(params, *a, **k)
                tap.init_params = [p.detach().numpy().copy() for p in params]
                tap.param_names = [n for n, _ in tap.model.named_parameters()]
                tap.init_buffers = {n: b.detach().numpy().copy() for n, b in tap.model.named_buffers()}

            def step(self, *a, **k):
                r = super().step(*a, **k)
                cur = tap.cur
                cur['params_after'] = [p.detach().numpy().copy() for g in self.param_groups for p in g['params']]
                cur['buffers_after'] = {n: b.detach().numpy().copy() for n, b in tap.model.named_buffers()}
                tap.steps.append(cur)
                tap._new_step()
                return r

        np.random.choice = choice
        np.random.rand = rand
        torch.distributions.normal._standard_normal = std_normal
        torch.nn.utils.clip_grad_norm_ = clip
        jm.optim.Adam = RecAdam
        return self

    def __exit__(self, *exc):
        jamie = import_reference()
        np.random.choice = self._orig['choice']
        np.random.rand = self._orig['rand']
        torch.distributions.normal._standard_normal = self._orig['std_normal']
        torch.nn.utils.clip_grad_norm_ = self._orig['clip']
        jamie.jamie.optim.Adam = self._orig['Adam']
        return False
