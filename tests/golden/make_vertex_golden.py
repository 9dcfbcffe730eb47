"""Golden vectors for the no-ghost vertex centring, produced by the UNMODIFIED reference
(yt/utilities/lib/interpolators.pyx:177-227 compiled into oracle/_ref by oracle/build_ref.py), called the way
grid_patch.py:284-296 calls it.  Run in the build container (needs /root/reference):

    python tests/golden/make_vertex_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

SHAPES = [(8, 8, 8), (5, 9, 3), (2, 2, 2), (16, 4, 7), (3, 2, 11)]


def cells_for(shape, k):
    """Inputs are regenerated from the seed by the tests; values straddle 0 so the log path sees NaNs."""
    return np.random.default_rng(1000 + k).random(shape) * 50.0 - 5.0


if __name__ == "__main__":
    from yt.utilities.lib.interpolators import ghost_zone_interpolate
    out = {}
    for k, shape in enumerate(SHAPES):
        cells = cells_for(shape, k)
        want = np.zeros(tuple(s + 1 for s in shape))
        ghost_zone_interpolate(1, cells, np.array([0.5] * 3), want, np.zeros(3))
        out["v%d" % k] = want
    np.savez_compressed(os.path.join(HERE, "vertex_golden.npz"), **out)
    print("wrote", len(out), "arrays")
