"""The four lenses beyond plane-parallel / perspective (lens.py:295-844): they only generate per-pixel rays, which the
samplers' GENERIC lens mode casts.  CPU: the ray sets (fisheye vectors against the compiled reference function,
geometry and projection round trips).  GPU: a scene rendered through each lens against the oracle fed the same rays."""
import os
import sys

import numpy as np
import pytest

import yt_b200 as yt
from oracle import oracle as O
from tests import common as T
from yt_b200.volume_rendering import lens as LN

LENSES = ["stereo-perspective", "fisheye", "spherical", "stereo-spherical"]


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("res,fov", [(16, 180.0), (33, 120.0), (64, 194.6)])
def test_fisheye_vectors_match_compiled_reference(res, fov):
    ref_dir = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from yt.utilities.lib.grid_traversal import arr_fisheye_vectors
    want = arr_fisheye_vectors(res, fov)
    got = LN.arr_fisheye_vectors(res, fov)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-15      # numpy vs libm sin/cos/atan2: last-bit differences


def _scene(lens_type, res=(48, 40)):
    arr = np.random.default_rng(9).random((12, 12, 12)) + 0.2
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=8)
    sc = yt.create_scene(ds, lens_type=lens_type)
    sc.camera.resolution = (res[0], res[0]) if lens_type == "fisheye" else res
    sc.camera.position = [1.9, 1.5, 1.2]
    if "spherical" in lens_type or lens_type == "fisheye":
        # these lenses look outwards from inside the volume.  (Not ON a cell face: a fisheye ray that runs exactly in the
        # face between two bricks makes the reference sample an out-of-range cell over a NEGATIVE-length segment
        # (grid_traversal.pyx:104-138 with cur == dims and tmax == -0.0) -- undefined input we walk without touching memory.)
        sc.camera.position = [0.45, 0.55, 0.52]
    return ds, sc


@pytest.mark.parametrize("lens_type", LENSES)
def test_ray_sets_have_the_reference_layout(lens_type):
    ds, sc = _scene(lens_type)
    cam = sc.camera
    src = sc.get_source()
    p = cam._get_sampler_params(src)
    nx, ny = cam.resolution
    assert p["lens_type"] == lens_type
    assert np.asarray(p["vp_pos"]).shape == (nx, ny, 3) and np.asarray(p["vp_dir"]).shape == (nx, ny, 3)
    assert p["image"].shape == (nx, ny, 4)
    d = np.asarray(p["vp_dir"])
    if lens_type == "spherical":
        n = np.linalg.norm(d, axis=2)
        assert np.allclose(n, n[0, 0])                      # all rays have the same length (max_length)
        assert np.allclose(np.asarray(p["vp_pos"]), cam.position)
    if lens_type == "stereo-spherical":
        assert np.allclose(d[:, : ny // 2], d[:, ny // 2:])   # both eyes share directions, origins differ
        assert not np.allclose(np.asarray(p["vp_pos"])[:, : ny // 2], np.asarray(p["vp_pos"])[:, ny // 2:])
    if lens_type == "stereo-perspective":
        left, right = np.asarray(p["vp_pos"])[: nx // 2], np.asarray(p["vp_pos"])[nx // 2:]
        sep = np.linalg.norm(left[0, 0] - right[0, 0])
        assert np.isclose(sep, 2 * abs(cam.lens.disparity), rtol=1e-2)   # (each eye's east vector is rotated by -/+ the disparity angle)
    if lens_type == "fisheye":
        corner = d[0, 0]
        assert np.all(corner == 0.0)                          # outside the dome: zero vectors (grid_traversal.pyx:340-342)


@pytest.mark.parametrize("lens_type", ["fisheye"])
def test_project_to_plane_inverts_the_ray_map(lens_type):
    # (SphericalLens.project_to_plane works in world axes while its rays are rotated into the camera frame, lens.py:655-725:
    # not an inverse pair in the reference either)
    ds, sc = _scene(lens_type, res=(64, 64))
    cam = sc.camera
    cam.lens.setup_box_properties(cam)
    p = cam._get_sampler_params(sc.get_source())
    d = np.asarray(p["vp_dir"])
    idx = [(20, 30), (40, 22), (32, 32), (12, 45)]
    pts = np.array([cam.position + 0.3 * d[i, j] / np.linalg.norm(d[i, j]) for i, j in idx])
    px, py, dz = cam.lens.project_to_plane(cam, pts)
    for (i, j), a, b in zip(idx, px, py):
        assert abs(a - i) <= 1 and abs(b - j) <= 1
    assert np.all(dz > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("lens_type", LENSES)
def test_scene_through_every_lens_matches_oracle(lens_type, gpu):
    from tests.test_scene_gpu import oracle_render
    ds, sc = _scene(lens_type)
    im = np.asarray(sc.render())
    want, nb = oracle_render(sc)
    assert nb == 8 and im.shape == want.shape
    assert T.normalised_diff(im, want, False) <= 1e-4
    assert im[:, :, :3].max() > 0
