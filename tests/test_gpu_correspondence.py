"""F estimation: ``compute_distances`` + ``Prime_Dual`` against outputs of the unmodified
reference (tests/golden/prime_dual.npz, written by tests/golden/make_prime_dual.py), and the default constructor path
``JAMIE().fit_transform(data)`` that uses them."""
import contextlib
import io
import os

import numpy as np
import pytest

from tests import parity_util as U
from tests.golden_util import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def test_prime_dual_matches_reference_output():
    from jamie_b200.correspondence import distance_function, prime_dual
    z = np.load(os.path.join(GOLDEN_DIR, 'prime_dual.npz'))
    d0, d1 = z['data0'], z['data1']
    fn = distance_function('euclidean', 40)
    np.testing.assert_allclose(fn(d0), z['dist0'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(fn(d1), z['dist1'], rtol=1e-12, atol=1e-12)
    for tag, epochs, tol in (('short', 60, 1e-4), ('long', 600, 2e-3)):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            F = prime_dual(z['dist0'], z['dist1'], d0.shape[1], d1.shape[1], epoch_pd=epochs, log_pd=100)
        ref = z[f'F_{tag}']
        assert F.shape == ref.shape == (48, 40) and F.dtype == np.float32
        err = U.rel(F, ref)
        print(tag, 'rel err vs reference F', err)
        assert err < tol, (tag, err)
        # the reference's progress lines (epoch / err / alpha), same numbers to the printed precision +- 1 digit
        ref_lines = [ln for ln in str(z[f'log_{tag}']).splitlines() if ln.startswith('epoch:[')]
        got_lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith('epoch:[')]
        assert len(got_lines) == len(ref_lines)
        for g, r in zip(got_lines, ref_lines):
            gv = [float(t.split(':')[1]) for t in g.split()[1:]]
            rv = [float(t.split(':')[1]) for t in r.split()[1:]]
            assert g.split()[0] == r.split()[0] and np.allclose(gv, rv, atol=2e-3), (g, r)


def test_default_constructor_path_estimates_F(capsys):
    """``use_f_tilde=True`` with no ``match_result`` (the reference's default): distances, Prime_Dual, then training with
    the F loss on the estimate -- no silent F = 0 downgrade."""
    from jamie import JAMIE
    rng = np.random.default_rng(2)
    t = rng.random(96)
    lat = np.stack([t, np.sin(4 * t), t ** 2], 1)
    data = [lat @ rng.normal(size=(3, 40)) + 0.02 * rng.normal(size=(96, 40)),
            lat @ rng.normal(size=(3, 28)) + 0.02 * rng.normal(size=(96, 28))]
    jm = JAMIE(output_dim=8, batch_size=96, pca_dim=None, epoch_pd=200, log_pd=100, epoch_DNN=60, min_epochs=20, log_DNN=30)
    emb = jm.fit_transform(dataset=data)
    out = capsys.readouterr().out
    for line in ('Shape of Raw data', 'Dataset 0: (96, 40)', 'Find correspondence between Dataset 1 and Dataset 2',
                 'epoch:[200/200] err:', 'Finished Matching!', 'Train coupled autoencoders', 'JAMIE Done!'):
        assert line in out, line
    F = jm.match_result[0]
    assert F.shape == (96, 96) and F.min() >= 0 and 0.5 < F.sum(1).mean() < 1.5      # soft row-stochastic estimate
    assert len(jm.dist) == 2 and jm.dist[0].shape == (96, 96)
    assert np.mean(jm.loss_history['F']) > 0                                            # the F term is live
    assert emb[0].shape == (96, 8) and np.all(np.isfinite(emb[0]))
    jm.engine.close()
