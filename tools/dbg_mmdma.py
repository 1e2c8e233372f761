import os, sys, json, io, contextlib
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from jamie import JAMIE
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'mmdma.npz'))
data1, data2 = z['data1'].astype(np.float64), z['data2'].astype(np.float64)
np.random.seed(42)
jm = JAMIE(min_epochs=6, epoch_DNN=int(sys.argv[1]) if len(sys.argv) > 1 else 8, pca_dim=None, use_f_tilde=False)
try:
    with contextlib.redirect_stdout(io.StringIO()):
        jm.fit_transform(dataset=[data1.copy(), data2.copy()], P=np.eye(data1.shape[0]))
except Exception as ex:
    print('EXC', ex)
h = jm.loss_history
for k in h:
    print(k, [float('%.4g' % v) for v in h[k][:8]])
