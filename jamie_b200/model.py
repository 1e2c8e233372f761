"""``edModelVar``: the state container / checkpoint shell of the coupled VAE.

Mirrors the module tree of the reference class (jamie/model.py:116-282) -- same child names and indices, same
parameter registration order, same attributes -- so that ``torch.save(model)`` files cross-load with the reference
(jamie/jamie.py:967-972) and user code that pokes ``jm.model.encoders[i]`` / ``fc_mus[i]`` / ``preprocessing[i]`` keeps
working.  All arithmetic of training and inference runs in the CUDA engine (``jamie_b200.engine.Engine``); the torch
modules here only hold the tensors for I/O.  ``forward`` / ``impute`` in eval mode route through the engine.
"""
import numpy as np
import torch
import torch.nn as nn

from .layout import bn_spec, param_spec
from .utilities import identity


class edModelVar(nn.Module):
    def __init__(self, input_dim, output_dim, preprocessing=None, preprocessing_inverse=None, sigma=None, dropout=None):
        super().__init__()
        self.num_modalities = len(input_dim)
        self.preprocessing = self.num_modalities * [identity] if preprocessing is None else preprocessing
        self.preprocessing_inverse = (
            self.num_modalities * [identity] if preprocessing_inverse is None else preprocessing_inverse)
        # jamie/model.py:144-145
        if dropout is None:
            dropout = .6 if max(input_dim) > 64 else 0

        def block(i, o):
            return [nn.Linear(i, o), nn.BatchNorm1d(o), nn.LeakyReLU(), nn.Dropout(dropout)]

        # construction order == reference order, so torch's global generator yields the same initial weights
        self.encoders = nn.ModuleList([
            nn.Sequential(*block(d, 2 * d), *block(2 * d, d)) for d in input_dim])
        self.fc_mus = nn.ModuleList([nn.Linear(d, output_dim) for d in input_dim])
        self.fc_vars = nn.ModuleList([nn.Linear(d, output_dim) for d in input_dim])
        self.decoders = nn.ModuleList([
            nn.Sequential(*block(output_dim, d), *block(d, 2 * d), nn.Linear(2 * d, d)) for d in input_dim])
        self.sigma = nn.Parameter(torch.rand(self.num_modalities))

    # ---- shape helpers
    @property
    def input_dims(self):
        return [self.encoders[i][0].in_features for i in range(self.num_modalities)]

    @property
    def output_dim(self):
        return self.fc_mus[0].out_features

    @property
    def dropout_p(self):
        return float(self.encoders[0][3].p)

    # ---- engine plumbing (never pickled)
    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop('_engine', None)
        return st

    def attach_engine(self, engine):
        object.__setattr__(self, '_engine', engine)
        from .utilities import set_gpu_projector
        set_gpu_projector(engine)   # preclass PCA projections of this process now run on the engine's GPU

    def engine(self):
        return self.__dict__.get('_engine', None)

    def packed_parameters(self):
        spec = param_spec(self.input_dims, self.output_dim)
        named = dict(self.named_parameters())
        assert [n for n, _ in spec] == list(named.keys()), 'parameter registration order differs from the reference'
        return [named[n].detach().cpu().numpy().astype(np.float32) for n, _ in spec]

    def packed_buffers(self):
        bufs = dict(self.named_buffers())
        return {k: v.detach().cpu().numpy() for k, v in bufs.items()}

    def push_to_engine(self):
        eng = self.engine()
        eng.set_params(self.packed_parameters())
        eng.set_bn_stats(self.packed_buffers())

    def pull_from_engine(self):
        eng = self.engine()
        named = dict(self.named_parameters())
        with torch.no_grad():
            for (n, _), t in zip(eng.spec, eng.get_params()):
                named[n].copy_(torch.from_numpy(t))
            bufs = dict(self.named_buffers())
            for k, v in eng.get_bn_stats().items():
                bufs[k].copy_(torch.from_numpy(np.asarray(v)))
        return self

    # ---- reference-shaped entry points (eval mode only; training lives in JAMIE.project_jamie)
    def _require_eval(self):
        if self.training:
            raise NotImplementedError(
                'jamie_b200 trains inside its CUDA engine (JAMIE.fit_transform); the module forward is eval-only')
        if self.engine() is None:
            raise RuntimeError('no CUDA engine attached to this model; use JAMIE.load_model / fit_transform')

    def forward(self, *X, corr=None):
        """Eval-mode forward: (zs, combined, X_hat, mus, logvars) with zs = mus (jamie/model.py:233-234). The reference's
        own eval callers discard everything but ``zs`` (jamie/jamie.py:798, 828); ``combined`` / ``X_hat`` are returned for
        the zero-correspondence case (combined = mus, X_hat = decode(mus)). A non-zero ``corr`` would mix the modalities
        in ``combine`` (jamie/model.py:245-259), which only the training step of the engine implements: it raises here
        rather than return tensors that silently ignore it."""
        self._require_eval()
        eng = self.engine()

        def host(x):
            return (x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)).astype(np.float32)

        if corr is not None and np.any(host(corr) != 0):
            raise NotImplementedError('edModelVar.forward in eval mode with a non-zero corr: combined / X_hat of the mixed '
                                      'latents are only computed inside the training step (jb_train_steps)')
        mus = [torch.as_tensor(eng.encode(i, host(x))) for i, x in enumerate(X)]
        xhat = [torch.as_tensor(eng.predict(i, i, host(x))) for i, x in enumerate(X)]
        return mus, mus, xhat, mus, None

    def impute(self, X, compose):
        self._require_eval()
        frm, to = compose
        x = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
        return torch.as_tensor(self.engine().predict(frm, to, x.astype(np.float32)))


def bn_prefixes(dims):
    return [p for p, _ in bn_spec(dims)]


# pickled by reference under the reference's module path
edModelVar.__module__ = 'jamie.model'
