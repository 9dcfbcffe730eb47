"""Edge cases of the cast through the C ABI, each against the oracle: empty brick lists, one-pixel images, rays
stopped before they start (z-buffer <= 0), bricks the view never reaches, NaN / inf data in a projection, a brick list
longer than one CTA-wide culling batch, and repeated casts into the same frame (the image accumulates, as in the
reference where every sampler(brick) call composites onto .aimage)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

pytestmark = pytest.mark.gpu


def _mods():
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    return S, G


def _unit_brick(n, seed, vertex=True):
    cells = np.random.RandomState(seed).uniform(0.05, 1.0, size=(n, n, n))
    data = T.vertex_centre(cells) if vertex else cells
    return ([data], np.ones((n, n, n), "uint8"), np.zeros(3), np.ones(3), [n] * 3)


def test_empty_brick_list_leaves_the_image_untouched(gpu):
    S, G = _mods()
    cam = T.make_camera(resolution=(20, 24))
    img0 = np.random.RandomState(0).uniform(size=(20, 24, 4))
    img = img0.copy()
    args, kw = T.sampler_args(cam, image=img)
    kw.pop("camera_data")
    s = S.VolumeRenderSampler(*args, T.colour_tf(), 10, **kw)
    assert s.aimage is img and np.array_equal(img, img0)          # nothing recorded: no work, no change
    b = _unit_brick(4, 1)
    bs = G.build_brickset([G.PartitionedGrid(0, b[0], b[1], b[2], b[3], np.array(b[4], "int64"))], s._ctx, s._storage)
    s.cast_bricks(bs, np.zeros(0, dtype=np.int64))                # a recorded but EMPTY order list
    assert np.array_equal(s.aimage, img0)


@pytest.mark.parametrize("kind", ["vr", "proj"])
def test_smallest_image(kind, gpu):
    """2 x 2 pixels: the plane-parallel ray generator divides by (nx - 1), (ny - 1) (lenses.pyx:186-187), so this is the
    smallest image with finite rays."""
    S, G = _mods()
    cam = T.make_camera(resolution=(2, 2), width=[0.3, 0.3, 1.5])
    sc = T.Scenario("2x2", kind, cam, [_unit_brick(6, 2, vertex=(kind == "vr"))], [0], tf=T.colour_tf(height=30.0))
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G)
    assert got.shape == (2, 2, 4) and want[..., 0].min() > 0
    if kind == "vr":
        assert T.normalised_diff(got, want, False) <= 1e-4
    else:
        assert np.abs(got[..., 0] - want[..., 0]).max() <= 1e-6 * np.abs(want[..., 0]).max()


def test_zbuffer_at_or_below_zero_stops_every_ray(gpu):
    S, G = _mods()
    cam = T.make_camera(resolution=(16, 16))
    zb = np.zeros((16, 16))
    zb[::2] = -3.0
    img0 = np.random.RandomState(3).uniform(size=(16, 16, 4))
    sc = T.Scenario("zb0", "vr", cam, [_unit_brick(6, 4)], [0], tf=T.colour_tf(height=30.0), zbuffer=zb, image0=img0)
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G)
    assert np.array_equal(want, img0)        # the reference never samples: walk_volume returns at intersect_t >= max_t
    assert np.array_equal(got, img0)


def test_nan_and_inf_cells_in_a_projection(gpu):
    S, G = _mods()
    cam = T.make_camera(resolution=(24, 24))
    b = _unit_brick(8, 5, vertex=False)
    b[0][0][2, 3, 4] = np.nan
    b[0][0][5, 5, 5] = np.inf
    sc = T.Scenario("nan_proj", "proj", cam, [b], [0])
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    ok = np.isfinite(want)
    assert ok.sum() > 100 and (~ok).sum() > 0
    assert np.abs(got[ok] - want[ok]).max() <= 1e-6 * np.abs(want[ok]).max()


def test_brick_list_longer_than_one_culling_batch(gpu):
    """600 thin slabs (> 2 batches of 256 list entries per CTA) stacked along x, cast in view order."""
    S, G = _mods()
    n = 600
    rng = np.random.RandomState(6)
    cam = T.make_camera(resolution=(40, 32), position=[1.4, 0.9, 1.2])
    bricks = []
    plane = rng.uniform(0.05, 1.0, size=(n + 1, 5, 5))
    for k in range(n):
        bricks.append(([np.ascontiguousarray(plane[k:k + 2])], np.ones((1, 4, 4), "uint8"), np.array([k / n, 0.0, 0.0]),
                       np.array([(k + 1) / n, 1.0, 1.0]), [1, 4, 4]))
    order = T.view_order([(b[2], b[3]) for b in bricks], cam.lens.viewpoint)
    sc = T.Scenario("slabs600", "vr", cam, bricks, order, tf=T.colour_tf(height=80.0))
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G)
    assert T.normalised_diff(got, want, False) <= 1e-4 and want[..., :3].max() > 0


def test_two_casts_accumulate_like_two_sampler_calls(gpu):
    """flush, then more bricks, then flush: the second cast composites onto the first's image."""
    S, G = _mods()
    cam = T.make_camera(resolution=(32, 32), position=[1.2, 0.8, 1.1])
    v = T.vertex_centre(np.random.RandomState(7).uniform(size=(8, 8, 8)))
    sub = T.split_brick([v], np.ones((8, 8, 8), "uint8"), np.zeros(3), np.ones(3), (2, 1, 1))
    order = T.view_order([(b[2], b[3]) for b in sub], cam.lens.viewpoint)
    sc = T.Scenario("two", "vr", cam, sub, order, tf=T.colour_tf(height=60.0))
    want = T.run_scenario(sc, O, O, num_threads=0)
    img = np.zeros((32, 32, 4))
    args, kw = T.sampler_args(cam, image=img)
    kw.pop("camera_data")
    s = S.VolumeRenderSampler(*args, sc.tf, 10, **kw)
    grids = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
             for i, b in enumerate(sub)]
    s(grids[order[0]])
    first = np.array(s.aimage, copy=True)
    s(grids[order[1]])
    both = np.array(s.aimage, copy=True)
    assert not np.array_equal(first, both)
    assert T.normalised_diff(both, want, False) <= 1e-4


@pytest.mark.parametrize("position", [[0.5, 0.5, 2.0], [0.5, 0.5, -1.0], [2.0, 0.5, 0.5], [-1.0, 0.5, 0.5], [0.5, 2.0, 0.5],
                                      [0.5, -1.0, 0.5], [1.7, 0.5, 1.9], [0.5, -0.8, -0.9]])
@pytest.mark.parametrize("masked", [False, True])
def test_colour_walk_axis_aligned_and_planar_views(position, masked, gpu):
    """The fused colour walk keeps a cell in ray-oriented corner order and mirrors the cell coordinates on descending
    axes; an axis the ray does not move along keeps its memory orientation and a fixed in-cell coordinate.  Views along
    +-x, +-y, +-z (two idle axes) and in a coordinate plane (one idle axis), with and without a cell mask."""
    S, G = _mods()
    n = 12
    north = [0.0, 1.0, 0.0] if abs(position[1] - 0.5) < 1e-12 else [0.0, 0.0, 1.0]
    cam = T.make_camera(resolution=(29, 31), position=position, focus=[0.5, 0.5, 0.5], width=[0.9, 0.9, 3.5], north=north)
    b = list(_unit_brick(n, 7))
    if masked:
        m = np.ones((n, n, n), "uint8")
        m[3:6, 2:9, 4:7] = 0
        m[::5, ::3, 1] = 0
        b[1] = m
    sc = T.Scenario("axis", "vr", cam, [tuple(b)], [0], tf=T.colour_tf(height=40.0))
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G)
    assert want[..., :3].max() > 0
    assert T.normalised_diff(got, want, False) <= 1e-4
    got64 = T.run_scenario(sc, S, G, sampler_kwargs={"storage": "f64"})
    assert T.normalised_diff(got64, want, False) <= 1e-4


def test_two_brick_sets_through_one_sampler_do_not_share_cached_footprints(gpu):
    """the frame caches the per-brick pixel footprints of the last (set, order).  Two different single-brick
    sets flushed through one sampler both have order [0], and the second temporary set may be allocated at the first
    one's address -- the cache must be keyed on the set's generation, not its pointer."""
    S, G = _mods()
    cam = T.make_camera(resolution=(64, 64), width=[1.5, 1.5, 1.5])
    img = np.zeros((64, 64, 4))
    args, kw = T.sampler_args(cam, image=img)
    kw.pop("camera_data")
    tf = T.colour_tf(height=60.0)
    s = S.VolumeRenderSampler(*args, tf, 10, **kw)
    want = np.zeros((64, 64, 4))
    a2, k2 = T.sampler_args(cam, image=want)
    k2.pop("camera_data")
    o = O.VolumeRenderSampler(*a2, tf, 10, **k2)
    # a small brick in one corner of the view, then a large one covering the whole view
    small = _unit_brick(4, 7)
    geoms = [(np.array([0.0, 0.0, 0.0]), np.array([0.2, 0.2, 0.2])), (np.zeros(3), np.ones(3))]
    for le, re in geoms:
        s(G.PartitionedGrid(0, small[0], small[1], le, re, np.array(small[4], "int64")))
        _ = s.aimage                                   # flush: a temporary brick set is built and freed
        o(O.PartitionedGrid(0, small[0], small[1], le, re, np.array(small[4], "int64")), num_threads=0)
    assert want[:, :, :3].max() > 0
    assert T.normalised_diff(np.array(s.aimage), want, False) <= 1e-4
    # the same set re-uploaded with another geometry and re-committed must not reuse the old footprints either
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    img[:] = 0.0
    want[:] = 0.0
    bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, s._storage)
    for le, re in geoms:
        bs.upload(0, small[0], small[1], le, re, small[4])
        bs.commit()
        s.cast_bricks(bs, np.zeros(1, dtype=np.int64))
        _ = s.aimage
        o(O.PartitionedGrid(0, small[0], small[1], le, re, np.array(small[4], "int64")), num_threads=0)
    assert T.normalised_diff(np.array(s.aimage), want, False) <= 1e-4
