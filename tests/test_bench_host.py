"""Host logic of bench.py: the kd slabs of the ranks and the sub-brick split of a slab that is too large
for 32-bit offsets.  The pieces of one reference brick must be met in increasing t by every ray
(SURVEY.md H14 / finding 6), whatever the view direction."""
import numpy as np
import pytest

import bench


def _entry_exit(lo, hi, o, d):
    t0, t1 = -np.inf, np.inf
    for a in range(3):
        if d[a] == 0.0:
            if not (lo[a] <= o[a] < hi[a]):
                return None
            continue
        ta, tb = (lo[a] - o[a]) / d[a], (hi[a] - o[a]) / d[a]
        t0, t1 = max(t0, min(ta, tb)), min(t1, max(ta, tb))
    return (t0, t1) if t0 < t1 else None


@pytest.mark.parametrize("vd", [(1.0, 1.0, 1.0), (-1.0, 0.3, 0.5), (0.2, -0.9, -0.1), (0.0, 0.0, 1.0), (-0.5, -0.5, -0.7)])
def test_split_slab_order_follows_the_rays(monkeypatch, vd):
    monkeypatch.setattr(bench, "MAX_BRICK_ELEMS", 5000)
    lo, hi = [0, 0, 0], [40, 24, 31]
    pieces = bench.split_slab(lo, hi, np.array(vd), 1)
    assert len(pieces) > 4
    # disjoint cover of the cell box, every piece small enough
    cover = np.zeros((40, 24, 31), dtype=np.int32)
    for plo, phi in pieces:
        assert np.prod([phi[a] - plo[a] + 1 for a in range(3)]) <= 5000
        cover[plo[0]:phi[0], plo[1]:phi[1], plo[2]:phi[2]] += 1
    assert cover.min() == 1 and cover.max() == 1
    rng = np.random.default_rng(0)
    d = np.array(vd)
    for _ in range(300):
        o = rng.uniform(-5, 45, 3) - 60.0 * d / np.linalg.norm(d)
        last = -np.inf
        for plo, phi in pieces:
            te = _entry_exit(plo, phi, o, d)
            if te is None:
                continue
            assert te[0] >= last - 1e-9, "piece met before an earlier one was left"
            last = te[1]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_kd_slabs_tile_the_grid(world):
    n = 64
    cover = np.zeros((n, n, n), dtype=np.int32)
    for r in range(world):
        lo, hi, path = bench.kd_slab(n, r, world)
        assert len(path) == int(np.log2(world))
        cover[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
    assert cover.min() == 1 and cover.max() == 1
