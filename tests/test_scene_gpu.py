"""Scene-level parity: yt_b200.create_scene(...).render() and off_axis_projection against the CPU
oracle driven brick by brick over the same kd-tree traversal (render_source.py:594-599)."""
import os

import numpy as np
import pytest

import yt_b200 as yt
from oracle import oracle as O
from tests import common as T
from yt_b200 import synthetic
from yt_b200.volume_rendering import utils as U

pytestmark = pytest.mark.gpu


def oracle_render(sc, camera=None):
    """What Scene.render does, with the oracle's samplers called once per brick in kd order."""
    camera = camera or sc.camera
    src = sc.get_source()
    src.zbuffer = None
    if not src._volume_valid:  # what @validate_volume does before the reference's render()
        fields, logs = [src.field], [src.log_field]
        src.volume.set_fields(fields, logs, no_ghost=(not src.use_ghost_zones))
        src._volume_valid = True
    params = camera._get_sampler_params(src)
    args = [np.atleast_3d(np.array(params["vp_pos"], dtype="float64")), np.atleast_3d(np.array(params["vp_dir"], dtype="float64")),
            np.zeros(3) if params["center"] is None else np.array(params["center"], dtype="float64"), params["bounds"],
            np.zeros(tuple(camera.resolution) + (4,)),
            np.array(params["x_vec"], dtype="float64"), np.array(params["y_vec"], dtype="float64"),
            np.array(params["width"], dtype="float64"), "KDTree"]
    kw = {"lens_type": params["lens_type"], "zbuffer": np.ones(tuple(camera.resolution))}
    if src.sampler_type == "volume-render":
        s = O.VolumeRenderSampler(*args, src.transfer_function, src.num_samples, camera_data=params["camera_data"], **kw)
    else:
        s = O.InterpolatedProjectionSampler(*args, src.num_samples, **kw)
    n = 0
    for b in src.volume.traverse(camera.lens.viewpoint):
        s(O.PartitionedGrid(b.parent_grid_id, b.my_data, b.source_mask, b.LeftEdge, b.RightEdge, b.dims), num_threads=0)
        n += 1
    im = np.array(s.aimage)
    if not src.transfer_function.grey_opacity:
        im[:, :, 3] = 1
    return np.rot90(im, k=2), n


def _amr_ds():
    g, dims = synthetic.nested_amr_grids(base_cells=16, tile=8, levels=3, seed=4)
    return yt.load_amr_grids(g, dims)


@pytest.mark.parametrize("lens", ["plane-parallel", "perspective"])
@pytest.mark.parametrize("grey", [False, True])
def test_create_scene_render_amr(lens, grey, gpu):
    ds = _amr_ds()
    sc = yt.create_scene(ds, lens_type=lens)
    sc.camera.resolution = (96, 80)
    if lens == "perspective":
        sc.camera.position = [1.7, 1.5, 1.3]
        sc.camera.width = [0.7, 0.7, 1.0]
    src = sc.get_source()
    src.tfh.grey_opacity = grey
    im = sc.render()
    assert im.shape == (96, 80, 4) and np.isfinite(im).all()
    want, nbricks = oracle_render(sc)
    assert nbricks > 20
    err = T.normalised_diff(np.asarray(im), want, grey)
    assert err <= 1e-4, err
    assert np.asarray(im)[:, :, :3].max() > 0
    if not grey:
        assert np.all(np.asarray(im)[:, :, 3] == 1.0)


def test_batched_equals_brick_by_brick(gpu):
    ds = _amr_ds()
    sc = yt.create_scene(ds)
    sc.camera.resolution = (64, 64)
    a = np.array(sc.render())
    sc.get_source().batched = False
    b = np.array(sc.render())
    assert np.allclose(a, b, rtol=0, atol=2e-6)


def test_uniform_grid_pieces_give_the_single_grid_image(gpu):
    """Finding 6 of SURVEY.md: cell-aligned sub-bricks reproduce the single-brick image."""
    arr = np.random.default_rng(7).random((24, 24, 24)) + 0.2
    ims = []
    for nprocs in (1, 8):
        ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=nprocs)
        sc = yt.create_scene(ds)
        sc.camera.resolution = (72, 72)
        src = sc.get_source()
        src.set_log(False)
        src.use_ghost_zones = False
        ims.append(np.array(sc.render()))
    # interior vertices differ between the two decompositions only through no-ghost extrapolation at
    # piece faces, so compare each with its own oracle instead of with each other
    for nprocs, im in zip((1, 8), ims):
        ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=nprocs)
        sc = yt.create_scene(ds)
        sc.camera.resolution = (72, 72)
        sc.get_source().set_log(False)
        want, _ = oracle_render(sc)
        assert T.normalised_diff(im, want, False) <= 1e-4


def test_projection_transfer_function_through_scene(gpu):
    from yt_b200.volume_rendering import ProjectionTransferFunction
    ds = _amr_ds()
    sc = yt.create_scene(ds)
    sc.camera.resolution = (64, 64)
    src = sc.get_source()
    src.set_log(False)  # no-ghost extrapolation goes negative at grid faces: log10 would put NaNs in both images
    src.set_transfer_function(ProjectionTransferFunction())
    im = np.asarray(sc.render())
    want, _ = oracle_render(sc)
    assert np.isfinite(want).all()
    scale = np.abs(want[:, :, 0]).max()
    err = np.abs(im[:, :, 0] - want[:, :, 0]).max() / scale
    assert err <= 1e-6, (err, scale, np.abs(im[..., 0]).max())


@pytest.mark.parametrize("weight", [None, "temperature"])
def test_off_axis_projection_matches_oracle(weight, gpu):
    g, dims = synthetic.nested_amr_grids(base_cells=16, tile=8, levels=2, seed=9)
    rng = np.random.default_rng(10)
    for spec in g:
        spec["temperature"] = rng.random(tuple(spec["dimensions"])) + 1.0
    ds = yt.load_amr_grids(g, dims)
    normal, center, width, res = [0.3, 0.5, 0.8], [0.5, 0.5, 0.5], 1.5, 64
    got = np.asarray(yt.off_axis_projection(ds, center, normal, width, res, "density", weight=weight))
    # oracle: the same loop with the oracle's ProjectionSampler
    from yt_b200.volume_rendering import Scene
    from yt_b200.volume_rendering.render_source import KDTreeVolumeSource
    sc = Scene()
    cam = sc.add_camera(ds)
    cam.set_width([width] * 3)
    cam.resolution = (res, res)
    n = np.array(normal) / np.linalg.norm(normal)
    cam.position = np.array(center) - width * n
    cam.focus = center
    vecs = np.identity(3)
    ax = np.cross(vecs, n).sum(axis=1).argmax()
    east = np.cross(vecs[ax, :], n).ravel()
    cam.switch_orientation(n, np.cross(n, east).ravel())
    import types
    p = cam._get_sampler_params(types.SimpleNamespace(zbuffer=None))
    img = np.zeros((res, res, 4))
    s = O.ProjectionSampler(np.atleast_3d(np.array(p["vp_pos"])), np.atleast_3d(np.array(p["vp_dir"])), np.array(p["center"]), p["bounds"], img,
                            np.array(p["x_vec"]), np.array(p["y_vec"]), np.array(p["width"]), "KDTree", 10, lens_type=p["lens_type"],
                            zbuffer=np.ones((res, res)))
    for grid, mask in ds.all_data().blocks:
        if weight is None:
            data = [np.ascontiguousarray(grid["density"] * mask)]
        else:
            data = [np.ascontiguousarray(grid["density"] * grid[weight] * mask), np.ascontiguousarray(grid[weight] * mask)]
        s(O.PartitionedGrid(grid.id, data, mask.astype("uint8"), grid.LeftEdge, grid.RightEdge, grid.ActiveDimensions.astype("int64")), num_threads=0)
    want = np.array(s.aimage)
    if weight is None:
        want = want[:, :, 0] * width
    else:
        nz = want[:, :, 1] != 0
        out = np.zeros((res, res))
        out[nz] = want[:, :, 0][nz] / want[:, :, 1][nz]
        want = out
    scale = np.abs(want).max()
    assert scale > 0
    assert np.abs(got - want).max() / scale <= 1e-6


@pytest.mark.parametrize("lens", ["plane-parallel", "perspective"])
@pytest.mark.parametrize("grey", [False, True])
def test_oversized_leaves_are_cut_into_pieces(lens, grey, gpu, monkeypatch):
    """A leaf with more elements than 32-bit offsets allow is cut into cell-aligned pieces that are cast in
    increasing t (amr_kdtree.split_box / view_order): same image as the uncut leaf and as the oracle."""
    from yt_b200 import amr_kdtree
    arr = np.random.default_rng(21).random((24, 20, 28)) + 0.2
    images = {}
    for limit in (2 ** 31, 1500):
        monkeypatch.setattr(amr_kdtree, "MAX_BRICK_ELEMS", limit)
        ds = yt.load_uniform_grid({"density": arr}, arr.shape)
        sc = yt.create_scene(ds, lens_type=lens)
        sc.camera.resolution = (80, 72)
        if lens == "perspective":
            sc.camera.position = [1.9, 1.3, 1.6]
            sc.camera.width = [0.8, 0.8, 1.0]
        sc.get_source().tfh.grey_opacity = grey
        images[limit] = np.array(sc.render())
        if limit == 1500:
            vol = sc.get_source().volume
            assert len(vol._piece_plan) == 1 and vol.device_bricks(sc.get_source().sampler._ctx, sc.get_source().sampler._storage).n >= 8
            want, _ = oracle_render(sc)
    assert T.normalised_diff(images[1500], images[2 ** 31], grey) <= 2e-5
    assert T.normalised_diff(images[1500], want, grey) <= 1e-4


def test_baseline_config0_fake_random_ds_scene(gpu):
    """BASELINE.json configs[0]: fake_random_ds(64), Scene + VolumeSource with the default transfer function (512 bins,
    10 layers, log10 density), default corner camera, PlaneParallelLens, 512 x 512 -- the reference's own CPU-runnable
    case (~1.3e7 cells, ~1.3e8 samples), against the oracle."""
    from yt_b200.testing import fake_random_ds
    ds = fake_random_ds(64)
    sc = yt.create_scene(ds)
    src = sc.get_source()
    assert tuple(sc.camera.resolution) == (512, 512) and sc.camera.lens.lens_type == "plane-parallel"
    assert src.log_field and src.num_samples == 10
    im = np.asarray(sc.render())
    assert src.transfer_function.nbins == 512
    want, nbricks = oracle_render(sc)
    assert nbricks == 1
    assert im.shape == (512, 512, 4) and want[:, :, :3].max() > 0
    assert T.normalised_diff(im, want, False) <= 1e-4
    assert np.all(im[:, :, 3] == 1.0)


def _orientation_scene(lens_type, res=400):
    """The scene of the reference's only golden-image test (tests/test_vr_orientation.py:36-60): the orientation toy
    volume seen from -x with z up, field of view twice the domain."""
    from yt_b200.testing import fake_vr_orientation_test_ds
    from yt_b200.volume_rendering import Scene, create_volume_source
    ds = fake_vr_orientation_test_ds()
    sc = Scene()
    sc.add_source(create_volume_source(ds, field=("gas", "density")))
    cam = sc.add_camera(ds, lens_type=lens_type)
    cam.resolution = (res, res)
    cam.position = np.array([-4.0, 0.0, 0.0])
    cam.switch_orientation(normal_vector=[1.0, 0.0, 0.0], north_vector=[0.0, 0.0, 1.0])
    cam.set_width(ds.domain_width * 2.0)
    return sc, cam


@pytest.mark.parametrize("view", ["perspective", "plane-parallel", "yaw", "pitch", "roll"])
def test_vr_orientation_views_match_the_oracle(view, gpu):
    """test_vr_orientation.py:62-94 as parity tests: both lens types, then yaw / pitch / roll by pi about the origin.
    Rays are axis-parallel, the domain is [-1, 1]^3 and most of the volume sits exactly on the lower table bound."""
    sc, cam = _orientation_scene("perspective" if view == "perspective" else "plane-parallel")
    if view in ("yaw", "pitch", "roll"):
        getattr(cam, view)(np.pi, rot_center=np.zeros(3))
    im = np.asarray(sc.render())
    want, nbricks = oracle_render(sc)
    assert nbricks == 1 and want[:, :, :3].max() > 0
    assert T.normalised_diff(im, want, False) <= 1e-4
    # handedness: the three +z cubes are brighter material above the image centre for the unrotated views
    assert np.isfinite(im).all()
