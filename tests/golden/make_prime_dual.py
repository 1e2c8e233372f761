"""Golden for the F estimation: the UNMODIFIED reference's ``compute_distances`` +
``Prime_Dual`` (jamie/jamie.py:839-890, 314-414) on a small pair, run in this container.

  python tests/golden/make_prime_dual.py

``distance_mode='euclidean'`` (sklearn pairwise distances inside the reference) because the default 'geodesic' mode calls
``unioncom.utils.geodesic_distances``, a third-party function that is not vendored (oracle/ref_harness.py).
Output: tests/golden/prime_dual.npz  (data, the reference's distance matrices and its F after epoch_pd iterations).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.ref_harness import import_reference  # noqa: E402


def main():
    jamie = import_reference()
    rng = np.random.default_rng(5)
    t = rng.random(48)
    lat = np.stack([t, np.sin(4 * t), t ** 2], 1)
    d0 = lat @ rng.normal(size=(3, 30)) + 0.02 * rng.normal(size=(48, 30))
    d1 = lat[:40] @ rng.normal(size=(3, 22)) + 0.02 * rng.normal(size=(40, 22))
    out = {}
    for tag, epochs in (('short', 60), ('long', 600)):
        jm = jamie.JAMIE(distance_mode='euclidean', epoch_pd=epochs, log_pd=100)
        jm.dataset = [d0, d1]
        jm.dataset_num = 2
        jm.row = [48, 40]
        jm.col = [30, 22]
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            jm.compute_distances(save_dist=True)
            F = jm.match()[0]
        out[f'F_{tag}'] = np.asarray(F, np.float32)
        out[f'log_{tag}'] = buf.getvalue()
        out['dist0'], out['dist1'] = jm.dist[0], jm.dist[1]
    kw = dict(epsilon=jm.epsilon, rho=jm.rho, delay=jm.delay)
    np.savez_compressed(os.path.join(HERE, 'prime_dual.npz'), data0=d0, data1=d1, kw=str(kw), **out)
    print(kw, out['F_long'].shape, float(out['F_long'].sum()), out['log_long'][-300:])


if __name__ == '__main__':
    main()
