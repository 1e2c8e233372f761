import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from jamie_b200.engine import Engine
from tests import parity_util as U
dims, L, B, p, n = [200, 100], 8, 64, 0.5, 256
data = U.synth_pair(n, dims, seed=1)
params = U.torch_like_init(dims, L, seed=2)
eng = Engine(dims, L, B, p)
eng.set_params(params)
for i in range(2):
    eng.set_dataset(i, data[i])
eng.set_prior_diag(np.ones(n, np.float32)); eng.set_f_dense(None)
rng = np.random.default_rng(0)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 3
idx = np.stack([rng.choice(n, B, replace=False) for _ in range(ns)])
eng.upload_plan(idx, idx, np.full(ns, 0.25))
eng.train_steps(ns)
print(eng.read_losses(ns))
