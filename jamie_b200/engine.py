"""Thin Python owner of one ``jb_engine`` handle (see ``include/jamie_b200.h``).

Everything numeric happens inside ``libjamie_b200.so``; this class only marshals numpy arrays / device pointers and
raises the library's error strings.  Packed parameter / BatchNorm orders are the reference's (``layout.py``).
"""
import ctypes as C

import numpy as np

from . import _lib
from .layout import bn_spec, param_spec


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, dims, latent, max_batch, dropout, lr=1e-3, loss_weights=None, pf_ratio=1.0, seed=666,
                 device=0, world_size=1, betas=(0.9, 0.999), adam_eps=1e-8, max_grad_norm=1.0):
        self.lib = _lib.load()
        cfg = _lib.JbConfig()
        cfg.dims[0], cfg.dims[1] = int(dims[0]), int(dims[1])
        cfg.latent = int(latent)
        cfg.max_batch = int(max_batch)
        cfg.dropout = float(dropout)
        cfg.lr = float(lr)
        cfg.beta1, cfg.beta2, cfg.adam_eps = float(betas[0]), float(betas[1]), float(adam_eps)
        cfg.max_grad_norm = float(max_grad_norm)
        lw = [1, 1, 1, 1] if loss_weights is None else list(loss_weights)
        assert len(lw) == 4, f'There are 4 losses and {len(lw)} weights'
        for k in range(4):
            cfg.loss_w[k] = float(lw[k])
        cfg.pf_ratio = float(pf_ratio)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.device = int(device)
        cfg.world_size = int(world_size)
        self.dims = [int(dims[0]), int(dims[1])]
        self.latent = int(latent)
        self.max_batch = int(max_batch)
        self.device = int(device)
        h = C.c_void_p()
        _lib.check(self.lib.jb_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.spec = param_spec(self.dims, self.latent)
        self.n_params = int(self.lib.jb_num_params(self.h))
        assert self.n_params == sum(int(np.prod(s)) for _, s in self.spec)
        self.n_bn = int(self.lib.jb_num_bn_floats(self.h))
        self.plan_steps = 0
        self.plan_batch = 0

    def close(self):
        if getattr(self, 'h', None):
            from . import utilities
            if utilities._GPU_PROJECTOR is self:
                utilities.set_gpu_projector(None)
            self.lib.jb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters / state ------------------------------------------------------------------------------------
    def set_params(self, tensors):
        """tensors: list of arrays in named_parameters() order (or one packed 1-D array)."""
        packed = self._pack(tensors)
        _lib.check(self.lib.jb_set_params(self.h, _ptr(packed), packed.size))

    def get_params(self, as_list=True):
        packed = np.empty(self.n_params, np.float32)
        _lib.check(self.lib.jb_get_params(self.h, _ptr(packed), packed.size))
        return self._unpack(packed) if as_list else packed

    def get_grads(self):
        packed = np.empty(self.n_params, np.float32)
        _lib.check(self.lib.jb_get_grads(self.h, _ptr(packed), packed.size))
        return self._unpack(packed)

    def get_adam_state(self):
        m = np.empty(self.n_params, np.float32)
        v = np.empty(self.n_params, np.float32)
        t = C.c_longlong(0)
        _lib.check(self.lib.jb_get_adam_state(self.h, _ptr(m), _ptr(v), m.size, C.byref(t)))
        return self._unpack(m), self._unpack(v), int(t.value)

    def set_adam_state(self, m, v, t):
        m = self._pack(m)
        v = self._pack(v)
        _lib.check(self.lib.jb_set_adam_state(self.h, _ptr(m), _ptr(v), m.size, int(t)))

    def _pack(self, tensors):
        if isinstance(tensors, np.ndarray) and tensors.ndim == 1 and tensors.size == self.n_params:
            return _f32(tensors)
        assert len(tensors) == len(self.spec), (len(tensors), len(self.spec))
        parts = []
        for (name, shp), t in zip(self.spec, tensors):
            t = np.asarray(t, np.float32)
            assert tuple(t.shape) == tuple(shp), (name, t.shape, shp)
            parts.append(t.reshape(-1))
        return np.ascontiguousarray(np.concatenate(parts))

    def _unpack(self, packed):
        out, o = [], 0
        for _, shp in self.spec:
            n = int(np.prod(shp))
            out.append(packed[o:o + n].reshape(shp).copy())
            o += n
        return out

    def set_bn_stats(self, buffers):
        """buffers: dict '<prefix>.running_mean' / '.running_var' / '.num_batches_tracked' (state_dict names)."""
        parts, nbt = [], (C.c_longlong * 8)()
        for k, (pre, w) in enumerate(bn_spec(self.dims)):
            rm = np.asarray(buffers[pre + '.running_mean'], np.float32)
            rv = np.asarray(buffers[pre + '.running_var'], np.float32)
            assert rm.shape == (w,) and rv.shape == (w,)
            parts += [rm, rv]
            nbt[k] = int(buffers.get(pre + '.num_batches_tracked', 0))
        packed = np.ascontiguousarray(np.concatenate(parts))
        _lib.check(self.lib.jb_set_bn_stats(self.h, _ptr(packed), packed.size, nbt))

    def get_bn_stats(self):
        packed = np.empty(self.n_bn, np.float32)
        nbt = (C.c_longlong * 8)()
        _lib.check(self.lib.jb_get_bn_stats(self.h, _ptr(packed), packed.size, nbt))
        out, o = {}, 0
        for k, (pre, w) in enumerate(bn_spec(self.dims)):
            out[pre + '.running_mean'] = packed[o:o + w].copy()
            out[pre + '.running_var'] = packed[o + w:o + 2 * w].copy()
            out[pre + '.num_batches_tracked'] = np.array(int(nbt[k]), np.int64)
            o += 2 * w
        return out

    # ---- data ---------------------------------------------------------------------------------------------------
    def set_dataset(self, mod, X, stream=0):
        """X: numpy [n, D_mod] (host) or a torch CUDA tensor (device, fp32, row-major)."""
        if hasattr(X, 'data_ptr'):
            assert X.is_cuda and X.dim() == 2 and X.shape[1] == self.dims[mod] and X.dtype.is_floating_point
            X = X.float().contiguous()
            _lib.check(self.lib.jb_set_dataset(self.h, mod, C.c_void_p(X.data_ptr()), X.shape[0], X.stride(0), 1,
                                               C.c_void_p(stream)))
            return
        X = _f32(X)
        assert X.ndim == 2 and X.shape[1] == self.dims[mod], (X.shape, self.dims)
        _lib.check(self.lib.jb_set_dataset(self.h, mod, _ptr(X), X.shape[0], X.shape[1], 0, C.c_void_p(stream)))

    def set_prior_diag(self, m):
        if m is None:
            _lib.check(self.lib.jb_set_prior_diag(self.h, None, 0))
        else:
            m = _f32(m)
            _lib.check(self.lib.jb_set_prior_diag(self.h, _ptr(m), m.size))

    def set_prior_dense(self, P):
        P = _f32(P)
        _lib.check(self.lib.jb_set_prior_dense(self.h, _ptr(P), P.shape[0], P.shape[1]))

    def set_f_dense(self, F):
        if F is None:
            _lib.check(self.lib.jb_set_f_dense(self.h, None, 0, 0))
        else:
            F = _f32(F)
            _lib.check(self.lib.jb_set_f_dense(self.h, _ptr(F), F.shape[0], F.shape[1]))

    # ---- training -----------------------------------------------------------------------------------------------
    def upload_plan(self, idx0, idx1, kl_anneal, stream=0):
        idx0 = np.ascontiguousarray(idx0, dtype=np.int64)
        idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
        assert idx0.ndim == 2 and idx0.shape == idx1.shape
        kl = np.ascontiguousarray(np.broadcast_to(np.asarray(kl_anneal, np.float64), (idx0.shape[0],)))
        _lib.check(self.lib.jb_upload_plan(self.h, _ptr(idx0), _ptr(idx1), _ptr(kl), idx0.shape[0], idx0.shape[1],
                                           C.c_void_p(stream)))
        self.plan_steps, self.plan_batch = idx0.shape

    def inject(self, eps, masks, stream=0):
        e0, e1 = _f32(eps[0]), _f32(eps[1])
        ms = [np.ascontiguousarray(m, dtype=np.uint8) for m in masks]
        assert len(ms) == 8
        arr = (C.c_void_p * 8)(*[m.ctypes.data for m in ms])
        _lib.check(self.lib.jb_inject_randomness(self.h, _ptr(e0), _ptr(e1), arr, C.c_void_p(stream)))

    def train_steps(self, n, stream=0):
        _lib.check(self.lib.jb_train_steps(self.h, int(n), C.c_void_p(stream)))

    def step_backward(self, stream=0):
        _lib.check(self.lib.jb_step_backward(self.h, C.c_void_p(stream)))

    def step_update(self, stream=0):
        _lib.check(self.lib.jb_step_update(self.h, C.c_void_p(stream)))

    def train_step_hostbatch(self, x0_ptr, x1_ptr, idx0, idx1, kl_anneal, stream=0):
        """One step with the batch rows coming from (pinned) host memory; returns the 8 loss scalars."""
        idx0 = np.ascontiguousarray(idx0, dtype=np.int64)
        idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
        out = np.empty(8, np.float32)
        _lib.check(self.lib.jb_train_step_hostbatch(self.h, C.c_void_p(x0_ptr), C.c_void_p(x1_ptr), _ptr(idx0), _ptr(idx1),
                                                    idx0.size, float(kl_anneal), _ptr(out), C.c_void_p(stream)))
        return out

    def hostbatch_submit(self, x0_ptr, x1_ptr, idx0, idx1, kl_anneal, stream=0):
        """Asynchronous train_step_hostbatch: returns at once; at most two steps in flight (jb_hostbatch_submit)."""
        idx0 = np.ascontiguousarray(idx0, dtype=np.int64)
        idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
        _lib.check(self.lib.jb_hostbatch_submit(self.h, C.c_void_p(x0_ptr), C.c_void_p(x1_ptr), _ptr(idx0), _ptr(idx1),
                                                idx0.size, float(kl_anneal), C.c_void_p(stream)))

    def hostbatch_wait(self):
        """The 8 loss scalars of the oldest step submitted with hostbatch_submit (blocks until it has finished)."""
        out = np.empty(8, np.float32)
        _lib.check(self.lib.jb_hostbatch_wait(self.h, _ptr(out)))
        return out

    def step_backward_hostbatch(self, x0_ptr, x1_ptr, idx0, idx1, kl_anneal, stream=0):
        idx0 = np.ascontiguousarray(idx0, dtype=np.int64)
        idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
        _lib.check(self.lib.jb_step_backward_hostbatch(self.h, C.c_void_p(x0_ptr), C.c_void_p(x1_ptr), _ptr(idx0),
                                                       _ptr(idx1), idx0.size, float(kl_anneal), C.c_void_p(stream)))

    def bench_stage(self, stage, iters, stream=0):
        us = C.c_float()
        fl = C.c_double()
        _lib.check(self.lib.jb_bench_stage(self.h, int(stage), int(iters), C.byref(us), C.byref(fl), C.c_void_p(stream)))
        return float(us.value), float(fl.value)

    def phase_names(self):
        return [self.lib.jb_phase_name(p).decode() for p in range(int(self.lib.jb_num_phases()))]

    def profile_step(self, iters=20, stream=0):
        """Average microseconds of every phase of one training step INSIDE the persistent kernel (jb_profile_step)."""
        out = np.zeros(64, np.float32)
        n = C.c_int()
        _lib.check(self.lib.jb_profile_step(self.h, int(iters), _ptr(out), 64, C.byref(n), C.c_void_p(stream)))
        return out[:n.value].copy()

    def pca_project(self, X, components, mean, m, sdev, stream=0):
        """((X - mean) @ components.T - m) / sdev on the GPU (jb_pca_project: error-compensated split GEMM, fp32-class):
        the PCA projection + scalar standardisation of ``preclass.transform`` (jamie/utilities.py:660-670). Host arrays
        in, host fp32 array out."""
        X = np.ascontiguousarray(X, np.float32)
        comp = np.ascontiguousarray(components, np.float32)
        mu = np.ascontiguousarray(mean, np.float32)
        n, d = X.shape
        k = comp.shape[0]
        assert comp.shape[1] == d and mu.shape[0] == d
        out = np.empty((n, k), np.float32)
        _lib.check(self.lib.jb_pca_project(self.h, _ptr(X), n, d, _ptr(comp), _ptr(mu), k, float(m), float(sdev), _ptr(out), 0,
                                           C.c_void_p(stream)))
        return out

    def pca_inverse(self, Z, components, mean, m, sdev, stream=0):
        """(Z * sdev + m) @ components + mean on the GPU (jb_pca_inverse): ``preclass.inverse_transform``
        (jamie/utilities.py:672-678)."""
        Z = np.ascontiguousarray(Z, np.float32)
        comp = np.ascontiguousarray(components, np.float32)
        mu = np.ascontiguousarray(mean, np.float32)
        n, k = Z.shape
        d = comp.shape[1]
        assert comp.shape[0] == k and mu.shape[0] == d
        out = np.empty((n, d), np.float32)
        _lib.check(self.lib.jb_pca_inverse(self.h, _ptr(Z), n, k, _ptr(comp), _ptr(mu), d, float(m), float(sdev), _ptr(out), 0,
                                           C.c_void_p(stream)))
        return out

    def pca_colsum(self, X, stream=0):
        """Column sums (float64) of a host [n, d] fp32 matrix on the GPU (jb_pca_colsum): first pass of the PCA fit."""
        X = np.ascontiguousarray(X, np.float32)
        n, d = X.shape
        out = np.empty(d, np.float64)
        _lib.check(self.lib.jb_pca_colsum(self.h, _ptr(X), n, d, out.ctypes.data_as(C.c_void_p), 0, C.c_void_p(stream)))
        return out

    def pca_gram(self, X, mean, stream=0):
        """(X - mean)^T (X - mean) as float64 [d, d] (jb_pca_gram: split tensor-core GEMM over row chunks, float64
        accumulation): second pass of the PCA fit (jamie/jamie.py:436-452)."""
        X = np.ascontiguousarray(X, np.float32)
        mean = np.ascontiguousarray(mean, np.float64)
        n, d = X.shape
        assert mean.shape[0] == d
        out = np.empty((d, d), np.float64)
        _lib.check(self.lib.jb_pca_gram(self.h, _ptr(X), n, d, mean.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 0,
                                        C.c_void_p(stream)))
        return out

    def profile_detail(self):
        """[phases, 4] of the last profile_step: total, longest CTA work, mean CTA work, barrier tail (us)."""
        n = int(self.lib.jb_num_phases())
        out = np.zeros(4 * n + 12 * 7, np.float32)
        _lib.check(self.lib.jb_profile_detail(self.h, _ptr(out), out.size))
        self.gemm_stamps = out[4 * n:].reshape(12, 7).copy()   # CTA 0 role stamps of the 12 GEMM phases (us)
        return out[:4 * n].reshape(n, 4)

    def set_dist_method(self, name):
        """sim_diff_func branch of the training loss (jamie/jamie.py:484-502): 'euclidean' or 'cosine'."""
        _lib.check(self.lib.jb_set_dist_method(self.h, {'euclidean': 0, 'cosine': 1}[name]))

    def set_grad_accumulate(self, flag):
        _lib.check(self.lib.jb_set_grad_accumulate(self.h, int(bool(flag))))

    def grad_buffer(self):
        """(device pointer, float count) of the flat gradient buffer (+ 8 loss scalars) for the DP all-reduce."""
        p = C.c_void_p()
        n = C.c_longlong()
        _lib.check(self.lib.jb_grad_buffer(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def set_grad_buffer(self, tensor):
        """Gradients go into a caller-owned CUDA float32 tensor (e.g. NVLink symmetric memory); None restores the engine's own."""
        if tensor is None:
            _lib.check(self.lib.jb_set_grad_buffer(self.h, None, 0))
            self._ext_grad = None
        else:
            _lib.check(self.lib.jb_set_grad_buffer(self.h, C.c_void_p(tensor.data_ptr()), C.c_longlong(tensor.numel())))
            self._ext_grad = tensor   # keep it alive

    def set_exchange(self, rank, world, grad_ptrs, flag_ptrs, multicast_ptr=0):
        """In-kernel gradient exchange over peer memory: device pointers of every rank's gradient buffer / scratch block
        (multicast_ptr: the NVSwitch multicast address of the gradient buffers, 0 = peer loads / stores)."""
        if not grad_ptrs:
            _lib.check(self.lib.jb_set_exchange(self.h, 0, 1, None, None, None))
            return
        ga = (C.c_void_p * world)(*[int(p) for p in grad_ptrs])
        fa = (C.c_void_p * world)(*[int(p) for p in flag_ptrs])
        _lib.check(self.lib.jb_set_exchange(self.h, int(rank), int(world), ga, fa, C.c_void_p(int(multicast_ptr) or None)))

    def exchange_scratch_bytes(self):
        return int(self.lib.jb_exchange_scratch_bytes())

    def grad_tensor(self):
        """The gradient buffer as a torch CUDA tensor view (no copy) -- what torch.distributed all-reduces."""
        import torch
        ptr, n = self.grad_buffer()

        class _Buf:
            __cuda_array_interface__ = {'shape': (n,), 'typestr': '<f4', 'data': (ptr, False), 'version': 3,
                                        'strides': None}
        return torch.as_tensor(_Buf(), device=f'cuda:{self.device}')

    def read_losses(self, n=None, stream=0):
        n = self.plan_steps if n is None else int(n)
        out = np.empty((n, 8), np.float32)
        _lib.check(self.lib.jb_read_losses(self.h, _ptr(out), n, C.c_void_p(stream)))
        return out

    # ---- eval ---------------------------------------------------------------------------------------------------
    def _eval(self, fn_name, X, width_out, head, stream):
        fn = getattr(self.lib, fn_name)
        if hasattr(X, 'data_ptr'):
            import torch
            assert X.is_cuda and X.dim() == 2 and X.dtype == torch.float32 and X.stride(1) == 1
            out = torch.empty((X.shape[0], width_out), dtype=torch.float32, device=X.device)
            if X.shape[0]:
                _lib.check(fn(self.h, *head, C.c_void_p(X.data_ptr()), X.shape[0], X.stride(0),
                              C.c_void_p(out.data_ptr()), out.stride(0), 1, C.c_void_p(stream)))
            return out
        X = _f32(X)
        out = np.empty((X.shape[0], width_out), np.float32)
        if X.shape[0]:
            _lib.check(fn(self.h, *head, _ptr(X), X.shape[0], X.shape[1], _ptr(out), width_out, 0, C.c_void_p(stream)))
        return out

    def encode(self, mod, X, stream=0):
        assert X.shape[1] == self.dims[mod], (X.shape, self.dims)
        return self._eval('jb_encode', X, self.latent, (int(mod),), stream)

    def predict(self, frm, to, X, stream=0):
        assert X.shape[1] == self.dims[frm], (X.shape, self.dims)
        return self._eval('jb_predict', X, self.dims[to], (int(frm), int(to)), stream)

    def predict_into(self, frm, to, x_ptr, n, ldx, out_ptr, ldo, on_device, stream=0):
        """Raw-pointer form (pinned host or device buffers) used by the benchmark."""
        _lib.check(self.lib.jb_predict(self.h, int(frm), int(to), C.c_void_p(x_ptr), int(n), int(ldx),
                                       C.c_void_p(out_ptr), int(ldo), int(on_device), C.c_void_p(stream)))

    # ---- debug --------------------------------------------------------------------------------------------------
    def debug_read(self, name, shape):
        out = np.empty(int(np.prod(shape)), np.float32)
        n = self.lib.jb_debug_read(self.h, name.encode(), _ptr(out), out.size)
        if n < 0:
            raise RuntimeError('jamie_b200: ' + self.lib.jb_last_error().decode())
        return out[:n].reshape(shape) if n == out.size else out[:n]

    def launch_count(self):
        return int(self.lib.jb_launch_count(self.h))
