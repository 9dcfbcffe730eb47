"""The slab-walk ProjectionSampler cast (yt_b200/csrc/proj_kernels.cuh) against the oracle over the cases its own
arithmetic distinguishes: every dominant axis and direction sign, rays stationary on one or two axes, views whose slope
is exactly one cell per slab, non-cubic cells, off-centre bricks, 1-4 fields, a z-buffer, bricks entered through side
faces, several bricks (plane-parallel and perspective) and bricks longer than one 256-slab re-seed chunk."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _cast(mod, grid_mod, cam, bricks, order, zbuffer=None, image0=None, **kw):
    nx, ny = cam.resolution
    img = np.zeros((nx, ny, 4)) if image0 is None else image0.copy()
    args, kwargs = T.sampler_args(cam, zbuffer=None if zbuffer is None else zbuffer.copy(), image=img)
    kwargs.pop("camera_data")
    kwargs.update(kw)
    s = mod.ProjectionSampler(*args, 10, **kwargs)
    grids = [grid_mod.PartitionedGrid(i, [T.zero_padded(d) for d in b[0]], b[1], np.array(b[2], "float64"), np.array(b[3], "float64"),
                                      np.array(b[4], "int64")) for i, b in enumerate(bricks)]
    for i in order:
        s(grids[i], **({"num_threads": 0} if mod is O else {}))
    return np.array(s.aimage, copy=True)


def _compare(cam, bricks, order, zbuffer=None, image0=None):
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    want = _cast(O, O, cam, bricks, order, zbuffer, image0)
    got = _cast(S, G, cam, bricks, order, zbuffer, image0)
    assert np.isfinite(got).all()
    scale = np.abs(want).max()
    assert scale > 0
    err = np.abs(got - want).max() / scale
    assert err <= RTOL, err
    big = np.abs(want) > 1e-3 * scale
    assert (np.abs(got[big] - want[big]) / np.abs(want[big])).max() <= RTOL
    return err


def _brick(dims, seed, le=(0, 0, 0), re=(1, 1, 1), nf=1):
    rng = np.random.RandomState(seed)
    return ([rng.uniform(0.1, 1.0 + k, size=tuple(dims)) for k in range(nf)], np.ones(tuple(dims), "uint8"), np.array(le, "float64"),
            np.array(re, "float64"), list(dims))


VIEWS = {
    # position relative to the focus (0.5, 0.5, 0.5): dominant axis and signs vary
    "x+": [2.0, 0.9, 0.8], "x-": [-1.2, 0.3, 0.65], "y+": [0.7, 2.1, 0.9], "y-": [0.4, -1.0, 0.2],
    "z+": [0.8, 0.7, 2.2], "z-": [0.35, 0.6, -1.3], "corner": None, "corner-": [-0.5, -0.5, -0.5],
    "xy-diagonal": [1.5, 1.5, 0.5], "xz-diagonal": [1.5, 0.5, 1.5], "near-axis": [0.5 + 1e-7, 0.5 - 3e-8, 2.0],
}


@pytest.mark.parametrize("view", sorted(VIEWS))
@pytest.mark.parametrize("nf", [1, 3])
def test_single_brick_every_dominant_axis(view, nf, gpu):
    cam = T.make_camera(resolution=(56, 44), position=VIEWS[view], width=[1.6, 1.4, 2.0])
    _compare(cam, [_brick((20, 17, 23), 3, nf=nf)], [0])


def test_axis_aligned_views_with_rays_on_cell_faces(gpu):
    # rays stationary on two axes, running exactly along cell faces and edges (grid_traversal.pyx tie / edge rules)
    for pos, north in (([0.5, 0.5, 2.0], None), ([2.0, 0.5, 0.5], [0, 0, 1]), ([0.5, -2.0, 0.5], [0, 0, 1])):
        cam = T.make_camera(resolution=(17, 17), position=pos, width=[1.0, 1.0, 3.0], north=north)
        _compare(cam, [_brick((16, 16, 16), 5)], [0])


def test_non_cubic_cells_off_centre_brick_four_fields_zbuffer_and_initial_image(gpu):
    cam = T.make_camera(resolution=(61, 50), position=[1.3, 0.2, 1.9], focus=[0.4, 0.55, 0.5], width=[1.2, 1.4, 2.5])
    rng = np.random.RandomState(8)
    zb = rng.uniform(-0.2, 1.3, size=(61, 50))
    img0 = rng.uniform(0, 3.0, size=(61, 50, 4))
    _compare(cam, [_brick((12, 40, 9), 9, le=(0.1, 0.0, 0.2), re=(0.9, 1.0, 0.7), nf=4)], [0], zbuffer=zb, image0=img0)


def test_long_brick_crosses_reseed_chunks(gpu):
    # 700 slabs along the dominant axis: the fixed-point state is re-seeded at slabs 256 and 512
    cam = T.make_camera(resolution=(40, 36), position=[3.0, 1.1, 0.9], focus=[0.5, 0.5, 0.5], width=[1.3, 1.3, 4.0])
    _compare(cam, [_brick((700, 6, 7), 11)], [0])
    cam = T.make_camera(resolution=(40, 36), position=[0.9, 1.1, -3.0], focus=[0.5, 0.5, 0.5], width=[1.3, 1.3, 4.0])
    _compare(cam, [_brick((5, 6, 600), 12)], [0])


@pytest.mark.parametrize("lens", ["plane-parallel", "perspective"])
def test_many_bricks(lens, gpu):
    cells = np.random.RandomState(21).uniform(0.2, 1.0, size=(24, 24, 24))
    sub = T.split_brick([cells[:24, :24, :24]], np.ones((24, 24, 24), "uint8"), np.zeros(3), np.ones(3), (3, 2, 4))
    # split_brick cuts vertex arrays (shared planes duplicated): cut cell arrays instead
    bricks = []
    for d, m, le, re, dims in sub:
        lo = np.rint(le * 24).astype(int)
        bricks.append(([np.ascontiguousarray(cells[lo[0]:lo[0] + dims[0], lo[1]:lo[1] + dims[1], lo[2]:lo[2] + dims[2]])], m, le, re, list(dims)))
    if lens == "perspective":
        cam = T.make_camera("perspective", resolution=(52, 48), position=[1.7, 1.4, 1.2], width=[0.7, 0.7, 1.0])
    else:
        cam = T.make_camera(resolution=(52, 48), position=[1.2, 0.8, 1.1])
    _compare(cam, bricks, list(range(len(bricks))))


def test_camera_inside_the_brick(gpu):
    cam = T.make_camera("perspective", resolution=(40, 40), position=[0.55, 0.45, 0.6], focus=[0.1, 0.2, 0.0], width=[0.8, 0.8, 0.5])
    _compare(cam, [_brick((18, 18, 18), 31)], [0])


@pytest.mark.parametrize("name", ["proj_offaxis_zbuf", "proj_corner", "proj_perspective"])
def test_banded_render_equals_the_in_place_render(gpu, name, monkeypatch):
    """ytb200_render stages a projection frame on page-locked arrays in row bands (copy-in, cast and copy-out of different
    bands overlap on three streams): bit-identical to the one-launch render that touches the host arrays in place, with and
    without a z-buffer, for odd image sizes (the last band is clipped) and with initial values in the image."""
    import numpy as np
    from oracle import oracle as O
    from tests import common as T
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    from yt_b200.partitioned_grid import build_brickset
    sc = [s for s in T.scenarios(small=False) if s.name == name][0]
    nx, ny = sc.cam.resolution
    assert nx >= 64                      # the banded route needs at least four tile rows
    rng = np.random.RandomState(11)
    image0 = rng.uniform(0, 2.0, size=(nx, ny, 4))
    sc.image0 = image0
    want = T.run_scenario(sc, O, O, num_threads=0)
    outs = []
    for banded in (True, False):
        if banded:
            monkeypatch.delenv("YTB200_NO_BANDS", raising=False)
        else:
            monkeypatch.setenv("YTB200_NO_BANDS", "1")
        img = D.pinned_empty((nx, ny, 4))
        img[:] = image0
        zb = None
        if sc.zbuffer is not None:
            zb = D.pinned_empty((nx, ny))
            zb[:] = sc.zbuffer
        args, kwargs = T.sampler_args(sc.cam, zbuffer=zb, image=img)
        kwargs.pop("camera_data")
        s = S.ProjectionSampler(*args, sc.n_samples, **kwargs)
        grids = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
                 for i, b in enumerate(sc.bricks)]
        bs = build_brickset(grids, s._ctx, s._storage)
        for _ in range(2):   # twice: the streams and events of the banded route are reused
            img[:] = image0
            s._frame.render(s.kind, bs, np.array(sc.order), img, zbuffer=zb, tf=None, n_samples=sc.n_samples)
        outs.append(np.array(img, copy=True))
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(outs[0] - want).max() <= 1e-6 * np.abs(want).max()
