"""Host-side mirror of the scene graph: loaders, vertex centring, default transfer function, wiring.
No GPU needed (nothing here casts a ray)."""
import os
import sys

import numpy as np
import pytest

import yt_b200 as yt
from oracle import oracle as O
from tests import common as T
from yt_b200 import loaders
from yt_b200.volume_rendering import ColorTransferFunction, ProjectionTransferFunction, Scene, ZBuffer
from yt_b200.volume_rendering.image_array import ImageArray
from yt_b200.volume_rendering.render_source import KDTreeVolumeSource


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape", [(8, 8, 8), (5, 9, 3), (2, 2, 2), (16, 4, 7)])
def test_vertex_centring_matches_reference_ghost_zone_interpolate(shape):
    ref_dir = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from yt.utilities.lib.interpolators import ghost_zone_interpolate
    cells = np.random.default_rng(5).random(shape)
    want = np.zeros(tuple(s + 1 for s in shape))
    ghost_zone_interpolate(1, cells, np.array([0.5] * 3), want, np.zeros(3))  # grid_patch.py:284-296
    got = loaders.vertex_centre_no_ghost(cells)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 4e-16 * np.abs(want).max() * 8


def test_load_uniform_grid_and_pieces():
    arr = np.random.default_rng(0).random((12, 8, 10))
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=6, bbox=np.array([[0, 3.0], [-1, 1], [0, 5]]))
    assert len(ds.index.grids) == 6
    assert sum(int(np.prod(g.ActiveDimensions)) for g in ds.index.grids) == arr.size
    total = np.zeros_like(arr)
    for g in ds.index.grids:
        lo = g.get_global_startindex()
        hi = lo + g.ActiveDimensions
        total[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += g["density"]
    assert np.array_equal(total, arr)
    assert ds.field_info[("stream", "density")].take_log
    assert ds._resolve(("gas", "density")) == ("stream", "density")
    lo, hi = ds.all_data().quantities["Extrema"]("density", non_zero=True)
    assert lo == arr.min() and hi == arr.max()


def test_amr_child_mask_and_blocks():
    from yt_b200 import synthetic
    g, dims = synthetic.nested_amr_grids(base_cells=16, tile=8, levels=2, seed=1)
    ds = yt.load_amr_grids(g, dims)
    covered = sum(int((1 - gr.child_mask).sum()) for gr in ds.index.grids if gr.Level == 0)
    assert covered == 8 ** 3  # the central half (per axis) of a 16^3 base is refined
    fine = [gr for gr in ds.index.grids if gr.Level == 1]
    assert all(gr.Parent is not None for gr in fine)
    assert all(gr.child_mask.all() for gr in fine)


def test_default_transfer_function_follows_the_helper_recipe():
    arr = np.random.default_rng(1).random((8, 8, 8)) + 0.5
    ds = yt.load_uniform_grid({"density": arr}, arr.shape)
    sc = yt.create_scene(ds)
    src = sc.get_source()
    assert isinstance(src, KDTreeVolumeSource) and src.log_field and src.num_samples == 10
    tf = src.transfer_function
    assert isinstance(tf, ColorTransferFunction) and tf.nbins == 512 and not tf.grey_opacity
    assert np.allclose(tf.x_bounds, np.log10([arr.min(), arr.max()]))
    assert len(tf.features) == 10
    # alpha table rescaled so that its mean is 2 (transfer_function_helper.py:112-114)
    assert abs(tf.alpha.y.mean() - 2.0) < 1e-12
    cam = sc.camera
    assert np.array_equal(cam.position, ds.domain_right_edge) and np.array_equal(cam.focus, ds.domain_center)
    assert np.allclose(cam.width, 1.5) and cam.resolution == (512, 512)
    assert cam.lens.lens_type == "plane-parallel"
    src.set_log(False)
    assert src._transfer_function is None and not src._volume_valid
    src.set_transfer_function(ProjectionTransferFunction())
    assert src.sampler_type == "projection"
    with pytest.raises(RuntimeError):
        src.set_transfer_function("viridis")


def test_zbuffer_merge_and_image_rescale(tmp_path):
    rng = np.random.default_rng(2)
    a = ZBuffer(rng.random((6, 5, 4)), rng.random((6, 5)))
    b = ZBuffer(rng.random((6, 5, 4)), rng.random((6, 5)))
    c = a + b
    near = a.z < b.z
    assert np.array_equal(c.rgba[near], a.rgba[near]) and np.array_equal(c.rgba[~near], b.rgba[~near])
    assert np.array_equal(c.z, np.minimum(a.z, b.z))
    im = ImageArray(rng.random((7, 9, 4)) * 3.0)
    r = im.rescale(inline=False)
    assert abs(np.asarray(r[:, :, :3]).sum(axis=2).max() - 1.0) < 1e-12 or np.asarray(r).max() == 1.0
    assert np.asarray(r[:, :, 3]).max() == 1.0
    fn = im.write_png(str(tmp_path / "x.png"), sigma_clip=2.0)
    data = open(fn, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and b"IEND" in data[-12:]


def test_scene_bookkeeping():
    arr = np.random.default_rng(3).random((4, 4, 4)) + 0.1
    ds = yt.load_uniform_grid({"density": arr}, arr.shape)
    sc = Scene()
    src = yt.create_volume_source(ds, "density")
    sc.add_source(src)
    sc.add_camera(ds, lens_type="perspective")
    assert sc.get_source(0) is src and sc["source_00"] is src
    assert list(k for k, _ in sc.transparent_sources) == ["source_00"] and list(sc.opaque_sources) == []
    assert sc.camera.lens.lens_type == "perspective"
    with pytest.raises(RuntimeError):
        sc.add_camera(ds, lens_type="no-such-lens")


def test_octree_cell_dataset_feeds_the_octree_source_like_the_reference():
    """load_octree_cells: the arrays OctreeRayTracing / OctreeVolumeSource.render read from the data source
    (octree_raytracing.py:19-41, render_source.py:667-680) -- unitary centres, dx = 2**-(level + 1), levels shifted by
    lvl_min = min_level + 1, and vertex data whose axes 0 and 2 set_fields swaps back to x-major."""
    import yt_b200 as yt
    from yt_b200.loaders import OctreeIndex
    ipos, lvl, centre, dx = T.random_octree(depth=5, max_level=4, p_refine=0.45, seed=3)
    cells = T.octree_cells(centre, dx, 5)
    vd = cells[:, :8].T.reshape(2, 2, 2, -1)                  # indexed (x, y, z, cell)
    ds = yt.load_octree_cells(centre, lvl - 1, {"density": vd})
    src = ds.all_data()
    assert isinstance(ds.index, OctreeIndex) and list(src.blocks) == []
    assert np.array_equal(src["dx"], dx) and np.array_equal(src["grid_level"], lvl - 1)
    assert np.array_equal(np.stack([src[k] for k in "xyz"], -1), centre)
    assert ds.min_level == 0 and ds.max_level == int(lvl.max()) - 1
    # what OctreeRayTracing.__init__ computes from them equals the tree inputs of the scenario
    lvl_min = ds.min_level + 1
    depth = lvl_min + ds.max_level + 1
    assert depth == 5
    assert np.array_equal(np.floor(centre * (1 << depth)).astype("int64"), ipos)
    assert np.array_equal(src["grid_level"] + lvl_min, lvl)
    # set_fields' layout fix (octree_raytracing.py:93-98) on the source's vertex data gives corner x*4 + y*2 + z
    data = src.vertex_centered_data("density").swapaxes(0, 2).reshape(8, -1)
    assert np.array_equal(data.T, cells[:, :8])
    lo, hi = src.quantities["Extrema"]("density", non_zero=True)
    assert lo == cells[:, :8].min() and hi == cells[:, :8].max()
    assert ds.field_info[("stream", "density")].take_log
    with pytest.raises(RuntimeError):
        yt.load_octree_cells(centre, lvl - 1, {"density": vd[:, :, :, :-1]})


def test_fake_random_ds_draws_like_the_reference_fixture():
    """yt/testing.py:345-436: one RandomState(0x4D3D3D3) stream, density drawn first, then the velocities."""
    from yt_b200.testing import fake_random_ds
    ds = fake_random_ds(16, nprocs=8)
    prng = np.random.RandomState(0x4D3D3D3)
    want = {f: prng.random_sample((16, 16, 16)) for f in ("density", "velocity_x", "velocity_y", "velocity_z")}
    assert len(ds.index.grids) == 8 and [k[1] for k in ds.field_list] == list(want)
    for g in ds.index.grids:
        lo = np.rint((g.LeftEdge - ds.domain_left_edge) * 16).astype(int)
        hi = lo + g.ActiveDimensions
        for f, v in want.items():
            assert np.array_equal(g[("stream", f)], v[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]])
    neg = fake_random_ds([4, 5, 6], peak_value=3.0, fields=("density", "velocity_x"), negative=[False, True])
    prng = np.random.RandomState(0x4D3D3D3)
    d0 = prng.random_sample((4, 5, 6)) * 3.0
    d1 = (prng.random_sample((4, 5, 6)) - 0.5) * 3.0
    g = neg.index.grids[0]
    assert np.array_equal(g[("stream", "density")], d0) and np.array_equal(g[("stream", "velocity_x")], d1)
    assert ds.default_field == ("stream", "density") and ds.field_info[("stream", "density")].take_log


def test_camera_view_sequences_and_auto_defaults():
    """camera.py:310-346, 598-760: iter_rotate / iter_move / iter_zoom step the camera like n calls of rotate /
    set_position / zoom; auto=True frames the box of the cell centres."""
    arr = np.random.default_rng(3).random((8, 8, 8)) + 0.1
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=8)
    sc = Scene()
    cam = sc.add_camera(ds)
    ref = sc.add_camera(ds)
    assert sc.get_camera() is ref
    sc.set_camera(cam)
    assert sc.camera is cam
    steps = list(cam.iter_rotate(0.6, 3))
    for _ in range(3):
        ref.rotate(0.2)
    assert steps == [0, 1, 2]
    assert np.allclose(cam.unit_vectors, ref.unit_vectors) and np.allclose(cam.position, ref.position)
    p0 = cam.position.copy()
    for _ in cam.iter_move([2.0, 3.0, 4.0], 4):
        pass
    assert np.allclose(cam.position, [2.0, 3.0, 4.0])
    for _ in cam.iter_move([4.0, 1.5, 8.0], 5, exponential=True):
        pass
    assert np.allclose(cam.position, [4.0, 1.5, 8.0]) and not np.allclose(p0, cam.position)
    assert np.allclose(cam.normal_vector, (cam.focus - cam.position) / np.linalg.norm(cam.focus - cam.position))
    w0 = cam.width.copy()
    assert len(list(cam.iter_zoom(8.0, 3))) == 3
    assert np.allclose(cam.width[:2], w0[:2] / 8.0) and cam.width[2] == w0[2]
    auto = sc.add_camera(ds, auto=True)
    # cell centres span [1/16, 15/16] on every axis: the diagonal of that box is the width
    assert np.allclose(auto.width, np.sqrt(3.0) * 14.0 / 16.0)
    assert np.allclose(auto.position, ds.domain_right_edge) and np.allclose(auto.focus, ds.domain_center)


@pytest.mark.skipif(not os.path.exists("/root/reference/yt/testing.py"), reason="needs the reference checkout")
def test_orientation_fixture_equals_the_reference_construction():
    """yt_b200.testing.fake_vr_orientation_test_ds against yt/testing.py:653-738 evaluated as written (the function body
    only needs numpy; `import yt` itself does not work here)."""
    from yt_b200.testing import fake_vr_orientation_test_ds
    src = open("/root/reference/yt/testing.py").read()
    a = src.index("def fake_vr_orientation_test_ds")
    body = src[a:src.index("\ndef ", a + 10)].replace(
        "from yt.loaders import load_uniform_grid", "load_uniform_grid = lambda data, shape, bbox=None: (data['density'][0], bbox)")
    ns = {"np": np}
    exec(body, ns)
    for kw in ({}, {"N": 40, "scale": 3.0}):
        want, bbox = ns["fake_vr_orientation_test_ds"](**kw)
        ds = fake_vr_orientation_test_ds(**kw)
        assert np.array_equal(ds.index.grids[0][("stream", "density")], want)
        assert np.array_equal(ds.domain_left_edge, bbox[:, 0]) and np.array_equal(ds.domain_right_edge, bbox[:, 1])


def test_parallel_assign_fills_like_numpy():
    """device.parallel_assign: slabs along axis 0 on a thread pool; strided views (one channel of an image), odd sizes,
    small arrays (serial path) and non-finite values."""
    from yt_b200.device import parallel_assign
    img = np.empty((1031, 517, 4))
    parallel_assign(img, 0.0)
    parallel_assign(img[:, :, 3], 1)
    assert (img[:, :, :3] == 0).all() and (img[:, :, 3] == 1).all()
    z = np.empty((2048, 2048))
    assert parallel_assign(z, np.inf) is z and np.isinf(z).all()
    small = np.empty((5, 7))
    parallel_assign(small, -2.5)
    assert (small == -2.5).all()
