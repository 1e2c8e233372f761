"""Shared helpers of the GPU parity tests: run the same seeded step through the CUDA engine (C ABI) and the numpy
oracle and compare every intermediate, loss, gradient and post-Adam parameter."""
import numpy as np

from oracle import jamie_oracle as O

PRE_BN_BIAS = {f'{k}.{i}.{l}.bias' for k in ('encoders', 'decoders') for i in (0, 1) for l in (0, 4)}


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def torch_like_init(dims, L, seed=0):
    """Deterministic parameter init with the reference's distributions (Linear U(+-1/sqrt(fan_in)), BN 1/0,
    sigma U(0,1)); the values only need to be shared by both sides."""
    rng = np.random.default_rng(seed)
    out = []
    for name, shp in O.param_spec(dims, L):
        if name == 'sigma':
            out.append(rng.random(2).astype(np.float32))
        elif name.split('.')[-2] in ('1', '5') and not name.startswith('fc_'):
            out.append(np.ones(shp, np.float32) if name.endswith('weight') else np.zeros(shp, np.float32))
        else:
            fan_in = shp[1] if len(shp) == 2 else None
            if fan_in is None:   # bias of the Linear whose weight came just before
                fan_in = out[-1].shape[1]
            b = 1.0 / np.sqrt(fan_in)
            out.append(rng.uniform(-b, b, size=shp).astype(np.float32))
    return out


def synth_pair(n, dims, seed=0, latent=8):
    rng = np.random.default_rng(seed)
    t = rng.normal(size=(n, latent))
    out = []
    for d in dims:
        A = rng.normal(size=(latent, latent))
        Bm = rng.normal(size=(latent, d)) / np.sqrt(latent)
        x = np.tanh(t @ A) @ Bm + 0.1 * rng.normal(size=(n, d))
        x = (x - x.mean()) / x.std()
        out.append(x.astype(np.float32))
    return out


def draw_randomness(B, dims, L, p, seed):
    rng = np.random.default_rng(seed)
    eps = [rng.normal(size=(B, L)).astype(np.float32) for _ in range(2)]
    widths = [2 * dims[0], dims[0], 2 * dims[1], dims[1], dims[0], 2 * dims[0], dims[1], 2 * dims[1]]
    masks = [(rng.random((B, w)) >= p).astype(np.uint8) for w in widths]
    return eps, masks


TAPS = ['x', 'y1_', 'h1_', 'y2_', 'h2_', 'mulv', 'z', 'c', 'xhat', 'dxhat', 'dg2_', 'dy4_', 'dg1_', 'dy3_', 'dc',
        'dmulv', 'dh2_', 'dy2_', 'dh1_', 'dy1_']


def oracle_taps(fw, model):
    """Oracle-side values of the forward taps the engine exposes."""
    out = {}
    for i in range(2):
        c1, c2 = fw['enc'][i]
        out[f'x{i}'] = fw['x'][i]
        out[f'h2_{i}'] = fw['h2'][i]
        out[f'mulv{i}'] = np.concatenate([fw['mu'][i], fw['lv'][i]], axis=1)
        out[f'z{i}'] = fw['z'][i]
        out[f'c{i}'] = fw['c'][i]
        out[f'xhat{i}'] = fw['xhat'][i]
        out[f'h1_{i}'] = c2[0]          # input of the second encoder Linear
        d1, d2, g2 = fw['dec'][i]
        out[f'g1_{i}'] = d2[0]
        out[f'g2_{i}'] = g2
    return out
