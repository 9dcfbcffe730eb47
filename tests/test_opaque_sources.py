"""Opaque sources: the zlines / zpoints painters against the compiled reference and the golden
vectors (CPU; the painters are host C++ like the reference's Cython), the source classes, and -- on the GPU -- a scene
whose volume cast starts from the painted colours and stops at the painted depths."""
import os
import sys

import numpy as np
import pytest

import yt_b200 as yt
from oracle import oracle as O
from tests import common as T
from tests.golden.make_zraster_golden import N_CASES, case
from yt_b200.volume_rendering import BoxSource, CoordinateVectorSource, GridSource, LineSource, PointSource, ZBuffer
from yt_b200.zraster import zlines, zpoints

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zraster_golden.npz")


@pytest.mark.parametrize("k", range(N_CASES))
def test_painters_match_golden_bit_for_bit(k):
    gold = np.load(GOLD)
    c = case(k)
    a, b = c["image"].copy(), c["z"].copy()
    zlines(a, b, c["xs"], c["ys"], c["zs"], c["colors"], c["ppc"], c["thick"], c["flip"], c["crop"])
    assert np.array_equal(a, gold["lines_img%d" % k], equal_nan=True) and np.array_equal(b, gold["lines_z%d" % k])
    a, b = c["image"].copy(), c["z"].copy()
    zpoints(a, b, c["xs"], c["ys"], c["zs"], c["colors"], c["radii"], c["ppc"])
    assert np.array_equal(a, gold["points_img%d" % k], equal_nan=True) and np.array_equal(b, gold["points_z%d" % k])


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_painters_match_compiled_reference_on_random_input():
    ref_dir = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from yt.utilities.lib._zraster_ref import zlines as ref_lines, zpoints as ref_points
    for k in range(100, 160):
        c = case(k)
        a, b = c["image"].copy(), c["z"].copy()
        ref_lines(a, b, c["xs"], c["ys"], c["zs"], c["colors"], c["ppc"], c["thick"], c["flip"], c["crop"])
        e, f = c["image"].copy(), c["z"].copy()
        zlines(e, f, c["xs"], c["ys"], c["zs"], c["colors"], c["ppc"], c["thick"], c["flip"], c["crop"])
        assert np.array_equal(a, e, equal_nan=True) and np.array_equal(b, f), k
        a, b = c["image"].copy(), c["z"].copy()
        ref_points(a, b, c["xs"], c["ys"], c["zs"], c["colors"], c["radii"], c["ppc"])
        e, f = c["image"].copy(), c["z"].copy()
        zpoints(e, f, c["xs"], c["ys"], c["zs"], c["colors"], c["radii"], c["ppc"])
        assert np.array_equal(a, e, equal_nan=True) and np.array_equal(b, f), k


def test_painter_argument_errors():
    img, z = np.zeros((8, 8, 4)), np.full((8, 8), np.inf)
    with pytest.raises(ValueError):
        zlines(img, z, [1, 2, 3], [1, 2, 3], [0.1, 0.2, 0.3], np.ones((2, 4)))       # odd number of points
    with pytest.raises(ValueError):
        zlines(img.astype("float32"), z, [1, 2], [1, 2], [0.1, 0.2], np.ones((1, 4)))


def _scene(lens="plane-parallel"):
    arr = np.random.default_rng(3).random((16, 16, 16)) + 0.2
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=8)
    sc = yt.create_scene(ds, lens_type=lens)
    sc.camera.resolution = (72, 64)
    if lens == "perspective":
        sc.camera.position = [1.8, 1.6, 1.4]
        sc.camera.width = [0.8, 0.8, 1.0]
    return ds, sc


@pytest.mark.parametrize("lens", ["plane-parallel", "perspective"])
def test_box_source_paints_twelve_edges(lens):
    ds, sc = _scene(lens)
    box = BoxSource(ds.domain_left_edge, ds.domain_right_edge, color=[1.0, 0.5, 0.25, 1.0])
    assert box.positions.shape == (24, 3) and box.color_stride == 24
    zb = box.render(sc.camera)
    assert isinstance(zb, ZBuffer) and zb.rgba.shape == (72, 64, 4)
    hit = np.isfinite(zb.z)
    assert 100 < hit.sum() < 72 * 64 / 2
    assert np.allclose(zb.rgba[hit], [1.0, 0.5, 0.25, 1.0])
    assert np.all(zb.rgba[~hit] == 0.0)


def test_grid_point_line_and_axes_sources():
    ds, sc = _scene()
    gs = GridSource(ds.all_data(), alpha=0.4)
    assert gs.positions.shape == (8 * 24, 3) and gs.colors.shape == (8, 4) and np.all(gs.colors[:, 3] == 0.4)
    zb = gs.render(sc.camera)
    assert (zb.rgba[..., 3] > 0).sum() > 200
    pts = PointSource(np.array([[0.5, 0.5, 0.5], [0.2, 0.8, 0.4]]), colors=np.array([[1, 0, 0, 1.0], [0, 1, 0, 1.0]]), radii=2)
    zb = pts.render(sc.camera, zbuffer=zb)
    assert (zb.rgba[..., 0] == 1.0).sum() >= 9
    ln = LineSource(np.array([[[0.1, 0.1, 0.1], [0.9, 0.9, 0.9]]]), colors=np.array([[0, 0, 1, 1.0]]))
    assert ln.positions.shape == (2, 3)
    ln.render(sc.camera, zbuffer=zb)
    ax = CoordinateVectorSource(thickness=2)
    ax.render(sc.camera, zbuffer=zb)
    assert np.isfinite(zb.rgba).all()


@pytest.mark.gpu
@pytest.mark.parametrize("lens", ["plane-parallel", "perspective"])
@pytest.mark.parametrize("grey", [False, True])
def test_scene_with_opaque_sources_matches_oracle(lens, grey, gpu):
    """Scene.composite: opaque sources fill the ZBuffer, the volume cast starts from its colours and stops at its
    depths (scene.py:611-622, utils.py:79-84, image_samplers.pyx:183-185)."""
    ds, sc = _scene(lens)
    src = sc.get_source()
    src.tfh.grey_opacity = grey
    sc.annotate_domain(ds, color=[0.9, 0.8, 0.1, 0.6])
    sc.add_source(PointSource(np.random.default_rng(1).random((40, 3)), colors=np.tile([0.2, 0.9, 0.9, 1.0], (40, 1)), radii=3))
    sc.annotate_axes()
    got = np.asarray(sc.render())
    # expected: the same opaque pass on the host, then the ORACLE's sampler brick by brick from that buffer
    cam = sc.camera
    empty = cam.lens.new_image(cam)
    opaque = ZBuffer(empty, np.full(empty.shape[:2], np.inf))
    for _, s in sc.opaque_sources:
        s.render(cam, zbuffer=opaque)
    assert np.isfinite(opaque.z).sum() > 300
    src.zbuffer = opaque
    params = cam._get_sampler_params(src)
    image0 = np.array(opaque.rgba, dtype="float64").reshape(cam.resolution[0], cam.resolution[1], 4).copy()
    args = [np.atleast_3d(np.array(params["vp_pos"], dtype="float64")), np.atleast_3d(np.array(params["vp_dir"], dtype="float64")),
            np.array(params["center"], dtype="float64"), params["bounds"], image0,
            np.array(params["x_vec"], dtype="float64"), np.array(params["y_vec"], dtype="float64"),
            np.array(params["width"], dtype="float64"), "KDTree"]
    s = O.VolumeRenderSampler(*args, src.transfer_function, src.num_samples, camera_data=params["camera_data"],
                              lens_type=params["lens_type"], zbuffer=np.array(opaque.z))
    for b in src.volume.traverse(cam.lens.viewpoint):
        s(O.PartitionedGrid(b.parent_grid_id, b.my_data, b.source_mask, b.LeftEdge, b.RightEdge, b.dims), num_threads=0)
    want = np.array(s.aimage)
    if not grey:
        want[:, :, 3] = 1
    want = np.rot90(want, k=2)
    assert got.shape == want.shape
    err = T.normalised_diff(got, want, grey)
    assert err <= 1e-4, err
    # the annotation is visible: pixels the volume never touches keep the painted colour
    assert np.abs(got - np.rot90(np.zeros_like(image0), k=2)).max() > 0
