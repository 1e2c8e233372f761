"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the real reference)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['diag_drop', 'rep_F', 'hybrid', 'zeros_unequal', 'pca', 'multibatch', 'cosine']


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
        self.meta = json.loads(str(self.z['meta']))
        self.kw = self.meta['kw']
        self.n_steps = self.meta['n_steps']
        self.param_names = self.meta['param_names']
        self.n_params = len(self.param_names)

    def __getitem__(self, k):
        return self.z[k]

    def has(self, k):
        return k in self.z.files

    def init_params(self):
        return [self.z[f'init/p{k}'] for k in range(self.n_params)]

    def buffers(self, prefix):
        out = {}
        for k in self.z.files:
            if k.startswith(prefix + '/b/'):
                out[k[len(prefix) + 3:]] = self.z[k]
        return out

    def masks(self, s):
        out = []
        for k in range(8):
            shp = tuple(self.z[f's{s}/maskshape{k}'])
            bits = np.unpackbits(self.z[f's{s}/mask{k}'])[:int(np.prod(shp))]
            out.append(bits.reshape(shp).astype(np.uint8))
        return out

    def eps(self, s):
        return [self.z[f's{s}/eps0'], self.z[f's{s}/eps1']]

    def choices(self, s):
        out = []
        k = 0
        while f's{s}/choice{k}' in self.z.files:
            out.append(self.z[f's{s}/choice{k}'])
            k += 1
        return out

    def grads(self, s):
        return [self.z[f's{s}/g{k}'] for k in range(self.n_params)]

    def params_after(self, s):
        return [self.z[f's{s}/p{k}'] for k in range(self.n_params)]

    def P_dense(self):
        n0, n1 = self.meta['n']
        if self.has('P'):
            return self.z['P']
        return np.eye(n0) if n0 == n1 else np.zeros((n0, n1))

    def F_dense(self):
        n0, n1 = self.meta['n']
        return self.z['F'] if self.has('F') else np.zeros((n0, n1))
