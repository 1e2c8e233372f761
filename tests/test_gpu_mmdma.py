"""End-to-end parity on BASELINE configs[0]: the MMD-MA branching-manifold simulation (the reference's own example data,
README.md:84-122) through ``JAMIE(min_epochs=500).fit_transform`` + ``modal_predict`` on the B200 engine, against the
metrics the UNMODIFIED reference produced on the same inputs (tests/golden/mmdma.npz, written by
tests/golden/make_mmdma.py in the build container: three seeds of the CPU reference, 25-37 minutes each).

Tolerance (stated, as north_star asks): training is stochastic (dropout masks, eps, batch order differ by RNG stream),
and early stopping makes single runs vary a lot (the reference's three seeds stop after 2480 .. 3897 epochs; ours after
1500 .. 6500, measured over 7 seeds: tools/mmdma_seeds.py). The test therefore runs THREE seeds, like the reference
fixture, and compares the MEDIAN of each metric with the reference's own seed-to-seed band widened by its width:
  FOSCTTM        reference 0.0087 .. 0.0139   ->  our median < 0.020   (measured per seed: 0.005 .. 0.021)
  LTA (k = 20)   reference 0.940 .. 0.950     ->  our median > 0.925
  LTA (k = 5)    reference 0.937 .. 0.957     ->  our median > 0.915
  imputation r   reference [0.910 .. 0.913, 0.960 .. 0.973] (mean per-feature Pearson r) -> our median > [0.895, 0.945]
  epochs run     reference 2480 .. 3897 (early stopping)    ->  every run of ours in 1000 .. 8000
"""
import contextlib
import io
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mmdma.npz')


def test_mmdma_end_to_end_matches_reference_band():
    from jamie import JAMIE
    from jamie_b200 import evaluation as E
    z = np.load(GOLD)
    runs = json.loads(str(z['runs']))
    data1, data2 = z['data1'].astype(np.float64), z['data2'].astype(np.float64)
    type1, type2 = z['type1'], z['type2']
    ref = {k: [r[k] for r in runs] for k in ('foscttm', 'lta', 'lta5', 'epochs')}
    ref_r = np.array([r['impute_r'] for r in runs])
    assert all(r['lta_k'] == E.default_k([type1, type2]) for r in runs)        # same default k as the reference
    res = []
    for seed in (42, 1, 2):
        np.random.seed(seed)
        jm = JAMIE(min_epochs=500, pca_dim=None, use_f_tilde=False, manual_seed=666 + seed)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            emb = jm.fit_transform(dataset=[data1.copy(), data2.copy()], P=np.eye(data1.shape[0]))
            fos = jm.test_closer(emb)
            lta, k_def = jm.test_LabelTA(emb, [type1, type2], return_k=True)
            lta5 = E.test_LabelTA(emb, [type1, type2], k=5)
            imp = [jm.modal_predict(data2, 1), jm.modal_predict(data1, 0)]          # README.md:110-111
        r = [E.mean_feature_r(imp[0], data1), E.mean_feature_r(imp[1], data2)]
        epochs = len(jm.loss_history['KL'])
        last = {k: v[-1] for k, v in jm.loss_history.items()}
        res.append(dict(fos=fos, lta=lta, lta5=lta5, r0=r[0], r1=r[1], epochs=epochs, last=last))
        print(f'ours (seed {seed}): foscttm {fos:.5f} lta(k={k_def}) {lta:.4f} lta5 {lta5:.4f} impute r {r[0]:.4f} {r[1]:.4f} epochs {epochs}')
        assert 'foscttm: ' in buf.getvalue() and 'Finished Mapping!' in buf.getvalue()
        assert emb[0].shape == (300, 32) and emb[1].shape == (300, 32)
        assert 1000 <= epochs <= 8000, epochs
        jm.engine.close()
    print(f'reference: foscttm {ref["foscttm"]} lta {ref["lta"]} lta5 {ref["lta5"]} r {ref_r.tolist()} epochs {ref["epochs"]}')
    med = {k: float(np.median([x[k] for x in res])) for k in ('fos', 'lta', 'lta5', 'r0', 'r1')}
    assert med['fos'] < 0.020, med
    assert med['lta'] > 0.925 and med['lta5'] > 0.915, med
    assert med['r0'] > 0.895 and med['r1'] > 0.945, med
    # late-training loss magnitudes (KL~0.19, Rec~0.29 with F; here the recorded reference rows)
    ref_last = {k: [x['final_losses'][k] for x in runs] for k in ('KL', 'Rec', 'CosSim')}
    for k in ('KL', 'Rec', 'CosSim'):
        m = float(np.median([x['last'][k] for x in res]))
        assert 0.5 * min(ref_last[k]) < m < 2.0 * max(ref_last[k]), (k, m, ref_last[k])
