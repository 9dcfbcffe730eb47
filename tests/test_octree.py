"""The octree path: Octree<int> traversal (_octree_raytracing.hpp) + cast_through_octree
(image_samplers.pyx:202-299).  CPU: the C restatement against the UNMODIFIED compiled reference (bit for bit) and
against golden images the reference produced; GPU: the CUDA cast against the oracle and the golden images."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_octree_images.npz")
SCENARIOS = T.octree_scenarios()


def load_ref_octree():
    O.load_ref()
    from yt.utilities.lib._octree_raytracing import _OctreeRayTracing
    return _OctreeRayTracing


def same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_matches_reference_bitwise(sc, threads):
    samplers, grid = O.load_ref()
    ref = T.run_octree_scenario(sc, samplers, grid, load_ref_octree(), num_threads=threads)
    got = T.run_octree_scenario(sc, O, O, O._OctreeRayTracing, num_threads=threads)
    assert same(ref, got), np.nanmax(np.abs(ref - got))
    base = np.zeros_like(ref) if sc.image0 is None else sc.image0
    assert (ref != base).any(axis=2).mean() > 0.3     # the rays do hit the tree


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_oracle_matches_golden(sc):
    gold = np.load(GOLDEN)[sc.name]
    got = T.run_octree_scenario(sc, O, O, O._OctreeRayTracing, num_threads=2)
    assert same(gold, got), np.nanmax(np.abs(gold - got))


def test_oracle_ray_list_uniform_tree():
    """A full level-2 tree is a 4x4x4 lattice: the cells and (t_in, t_out) of a ray along x are known in closed form,
    and a ray pointing the other way meets the same cells in reverse order.  (Not EXACTLY along x: with a zero
    direction component the parametric traversal divides by zero and its midpoints are NaN -- in the reference too.)"""
    tree = O._OctreeRayTracing(np.zeros(3), np.ones(3), 3)
    g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3)
    tree.add_nodes((g * 2 + 1).astype("int32"), np.full(len(g), 2, "int32"), np.arange(len(g), dtype="int32"))
    keys, t = tree.cast_ray([-0.5, 0.3, 0.6], [2.0, 2e-6, 4e-6])
    assert list(keys) == [(i * 4 + 1) * 4 + 2 for i in range(4)]
    assert np.allclose(t, [[0.5 + 0.25 * i, 0.75 + 0.25 * i] for i in range(4)])
    keys_r, t_r = tree.cast_ray([1.5, 0.3, 0.6], [-1.0, -1e-6, 2e-6])
    assert list(keys_r) == list(keys[::-1])
    assert np.allclose(t_r, t)
    assert len(tree.cast_ray([-0.5, 1.3, 0.6], [1.0, 1e-6, 1e-6])[0]) == 0     # misses the domain


@pytest.mark.gpu
@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
@pytest.mark.parametrize("precision", ["fast", "exact"])
def test_cuda_matches_oracle_and_golden(sc, precision, gpu):
    """EXACT: everything fp64 in the reference's order, agreement to rounding.  FAST (the default): the volume render
    samples in fp32 and must meet the 1e-4 bar on the normalised RGBA; the projection kinds are fp64 either way."""
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    from yt_b200.octree_raytracing import _OctreeRayTracing
    want = T.run_octree_scenario(sc, O, O, O._OctreeRayTracing, num_threads=0)
    kw = {"precision": "exact", "storage": "f64"} if precision == "exact" else {}
    got = T.run_octree_scenario(sc, S, G, _OctreeRayTracing, sampler_kwargs=kw)
    gold = np.load(GOLDEN)[sc.name]
    for ref in (want, gold):
        assert np.array_equal(np.isnan(ref), np.isnan(got))
        if sc.kind == "vr":
            assert T.normalised_diff(got, ref, sc.grey) <= (1e-10 if precision == "exact" else 1e-4)
            if not sc.grey:
                assert np.array_equal(got[..., 3], ref[..., 3])      # alpha untouched, bit for bit
        else:
            scale = np.nanmax(np.abs(ref))
            assert np.nanmax(np.abs(got - ref)) <= 1e-12 * scale


@pytest.mark.gpu
def test_cuda_octree_errors(gpu):
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    from yt_b200.octree_raytracing import OctreeRayTracing, _OctreeRayTracing
    from yt_b200._lib import YTB200Error
    sc = SCENARIOS[0]
    with pytest.raises(YTB200Error):
        _OctreeRayTracing(np.zeros(3), np.ones(3), 40)
    tree = _OctreeRayTracing(np.zeros(3), np.ones(3), 3)
    with pytest.raises(ValueError):
        tree.add_nodes(np.zeros((1, 3), "int64"), np.zeros(1, "int32"), np.zeros(1, "int32"))
    with pytest.raises(YTB200Error):
        tree.add_nodes(np.zeros((1, 3), "int32"), np.full(1, 4, "int32"), np.zeros(1, "int32"))   # level > depth
    # plane-parallel description: the reference reads vp_pos out of bounds, here it is refused
    cam = T.make_camera("plane-parallel", (16, 16))
    args, kwargs = T.sampler_args(cam)
    args[8] = "Octree"
    kwargs.pop("camera_data")
    s = S.ProjectionSampler(*args, 10, **kwargs)
    tree.add_nodes(np.array([[2, 2, 2]], "int32"), np.array([1], "int32"), np.array([0], "int32"))
    pg = G.PartitionedGrid(0, np.zeros((1, 1, 14, 1)), np.ones((1, 14, 1), "uint8"), np.zeros(3), np.ones(3), np.ones(3, "int64"), n_fields=1)
    with pytest.raises(ValueError):
        s(pg, oct=tree)
    # the array constructor of octree_raytracing.py builds the same tree as explicit add_nodes
    ipos, lvl, centre, dx = T.random_octree(depth=5, max_level=3, p_refine=0.5, seed=1)
    o = OctreeRayTracing(centre, lvl - 1, 0, 3)      # lvl_min = 1, depth = 1 + 3 + 1 = 5
    t2 = _OctreeRayTracing(np.zeros(3), np.ones(3), 5)
    t2.add_nodes(ipos, lvl, np.arange(len(lvl), dtype=np.int32))
    assert o.octree.depth == 5 and o.octree.node_count == t2.node_count


@pytest.mark.gpu
@pytest.mark.parametrize("grey", [False, True])
def test_create_scene_octree_dataset(grey, gpu):
    """yt.create_scene on an octree dataset -> OctreeVolumeSource.render (render_source.py:619-693), against the oracle
    driven with the same sampler arguments."""
    import yt_b200 as yt
    from yt_b200.volume_rendering.render_source import OctreeVolumeSource
    ipos, lvl, centre, dx = T.random_octree(depth=5, max_level=4, p_refine=0.45, seed=3)
    cells = T.octree_cells(centre, dx, 5)
    ds = yt.load_octree_cells(centre, lvl - 1, {"density": cells[:, :8].T.reshape(2, 2, 2, -1)})
    with pytest.raises(ValueError):
        yt.create_scene(ds).render()          # plane-parallel: no per-pixel rays (the reference reads out of bounds)
    sc = yt.create_scene(ds, lens_type="perspective")
    src = sc.get_source()
    assert isinstance(src, OctreeVolumeSource) and src.use_ghost_zones and src.log_field
    sc.camera.resolution = (48, 40)
    sc.camera.position = [1.6, 1.4, 1.2]
    sc.camera.width = [0.8, 0.8, 1.0]
    src.tfh.grey_opacity = grey
    im = np.array(sc.render())
    # the oracle, with the arguments the scene built
    params = sc.camera._get_sampler_params(src)
    args = [np.atleast_3d(np.array(params["vp_pos"], dtype="float64")), np.atleast_3d(np.array(params["vp_dir"], dtype="float64")),
            np.array(params["center"], dtype="float64"), params["bounds"], np.zeros(tuple(sc.camera.resolution) + (4,)),
            np.array(params["x_vec"], dtype="float64"), np.array(params["y_vec"], dtype="float64"),
            np.array(params["width"], dtype="float64"), "Octree"]
    s = O.VolumeRenderSampler(*args, src.transfer_function, src.num_samples, camera_data=params["camera_data"],
                              lens_type=params["lens_type"], zbuffer=np.ones(tuple(sc.camera.resolution)))
    tree = O._OctreeRayTracing(np.zeros(3), np.ones(3), src.volume.octree.depth)
    tree.add_nodes(ipos, lvl, np.arange(len(lvl), dtype=np.int32))
    assert src.volume.octree.depth == 5
    logc = cells.copy()
    logc[:, :8] = np.log10(cells[:, :8])
    pg = O.PartitionedGrid(0, logc.reshape(1, -1, 14, 1), np.ones((len(lvl), 14, 1), "uint8"), np.zeros(3), np.ones(3),
                           np.ones(3, "int64"), n_fields=1)
    s(pg, oct=tree)
    want = np.array(s.aimage)
    if not grey:
        want[:, :, 3] = 1
    want = np.rot90(want, k=2)
    assert (want[:, :, :3] > 0).any(axis=2).mean() > 0.3
    assert T.normalised_diff(im, want, grey) <= 1e-4       # FAST precision, the scene default
