"""Golden vectors for the opaque-source painters, produced by the UNMODIFIED reference functions zlines / zpoints
(yt/utilities/lib/misc_utilities.pyx:244-404, sliced out and compiled by oracle/build_ref.py).  Run in the build
container (needs /root/reference):  python tests/golden/make_zraster_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

N_CASES = 8


def case(k):
    """Inputs of case k (regenerated from the seed by the tests)."""
    rng = np.random.default_rng(4242 + k)
    nx, ny = int(rng.integers(12, 40)), int(rng.integers(12, 40))
    n = 2 * int(rng.integers(2, 10))
    xs = rng.integers(-8, nx + 8, n).astype("int64")
    ys = rng.integers(-8, ny + 8, n).astype("int64")
    zs = rng.uniform(0, 1, n)
    ppc = int(rng.integers(1, 3))
    colors = rng.uniform(0, 1, (n, 4))
    if k % 3 == 0:
        colors[:, 3] = 1.0
    radii = rng.integers(0, 4, n).astype("int64")
    image = rng.uniform(0, 1, (nx, ny, 4)) * (k % 2)
    z = np.where(rng.uniform(size=(nx, ny)) < 0.5, np.inf, rng.uniform(0, 1, (nx, ny)))
    return dict(xs=xs, ys=ys, zs=zs, ppc=ppc, colors=colors, radii=radii, image=image, z=z,
                thick=1 + k % 3, flip=int(k % 4 == 0), crop=int(k % 5 == 0))


if __name__ == "__main__":
    from yt.utilities.lib._zraster_ref import zlines, zpoints
    out = {}
    for k in range(N_CASES):
        c = case(k)
        a, b = c["image"].copy(), c["z"].copy()
        zlines(a, b, c["xs"], c["ys"], c["zs"], c["colors"], c["ppc"], c["thick"], c["flip"], c["crop"])
        out["lines_img%d" % k], out["lines_z%d" % k] = a, b
        a, b = c["image"].copy(), c["z"].copy()
        zpoints(a, b, c["xs"], c["ys"], c["zs"], c["colors"], c["radii"], c["ppc"])
        out["points_img%d" % k], out["points_z%d" % k] = a, b
    np.savez_compressed(os.path.join(HERE, "zraster_golden.npz"), **out)
    print("wrote", len(out), "arrays")
