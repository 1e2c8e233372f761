"""MMD-MA end-to-end metrics over several numpy seeds (the band the GPU test asserts against): python tools/mmdma_seeds.py 5"""
import os, sys, io, json, contextlib
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from jamie import JAMIE
from jamie_b200 import evaluation as E
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'mmdma.npz'))
data1, data2 = z['data1'].astype(np.float64), z['data2'].astype(np.float64)
type1, type2 = z['type1'], z['type2']
runs = json.loads(str(z['runs']))
print('reference:', [(r['epochs'], round(r['foscttm'], 4), r['lta']) for r in runs])
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    np.random.seed(seed)
    jm = JAMIE(min_epochs=500, pca_dim=None, use_f_tilde=False, manual_seed=666 + seed)
    with contextlib.redirect_stdout(io.StringIO()):
        emb = jm.fit_transform(dataset=[data1.copy(), data2.copy()], P=np.eye(data1.shape[0]))
        fos = jm.test_closer(emb)
        lta = jm.test_LabelTA(emb, [type1, type2])
    h = jm.loss_history
    print(f'seed {seed}: epochs {len(h["KL"])} foscttm {fos:.4f} lta {lta:.3f} final KL {h["KL"][-1]:.3f} Rec {h["Rec"][-1]:.3f} CosSim {h["CosSim"][-1]:.3f}')
    jm.engine.close()
