"""Ghost-zone vertex centring (grid.get_vertex_centered_data(no_ghost=False), grid_patch.py:298-310 on top of the
smoothed covering grid, construction_data_containers.py:1398-1653).

The two compiled pieces of that path -- ``ghost_zone_interpolate`` (interpolators.pyx:177-227) and ``fill_region``
(misc_utilities.pyx:707-789) -- are checked bit for bit against the reference built into oracle/_ref.  The level loop
around them needs yt's data objects (region selectors, unit registry: not installable offline), so it is restated in
yt_b200/loaders.py and checked through its defining properties: on one level it is the periodic global array; ghost
zones covered by same-level grids hold those grids' cells; everything else is the trilinear interpolation of the next
coarser level's box; cells equal to zero never replace an interpolated value.  The device path must then equal the
host path bit for bit (fp64 storage)."""
import numpy as np
import pytest

import yt_b200 as yt
from yt_b200 import synthetic
from yt_b200 import loaders


def _ref_lib():
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref is not built")
    O.load_ref()
    from yt.utilities.lib import _fillregion_ref, interpolators
    return interpolators, _fillregion_ref


@pytest.mark.parametrize("case", [(1, (5, 6, 7), [0.5] * 3, (6, 7, 8), [0.0] * 3),
                                  (2, (6, 5, 9), [3.0, 7.0, -1.0], (10, 8, 15), [2.5, 8.5, -2.5]),
                                  (2, (4, 4, 4), [1.0, 1.0, 1.0], (6, 6, 6), [0.5, 0.5, 0.5])])
def test_ghost_zone_interpolate_equals_the_reference(case):
    interp, _ = _ref_lib()
    rf, ishape, ileft, oshape, oleft = case
    f = np.random.default_rng(3).random(ishape) * 7 - 2
    want = np.zeros(oshape)
    interp.ghost_zone_interpolate(rf, f, np.array(ileft, "float64"), want, np.array(oleft, "float64"))
    got = loaders.ghost_zone_interpolate(float(rf), f, ileft, oshape, oleft)
    assert np.array_equal(got, want)


def _nested(seed=2, zero_cells=False):
    grids, dims = synthetic.nested_amr_grids(base_cells=8, tile=4, levels=3, seed=seed)
    if zero_cells:
        grids[-1]["density"][1:3, 2, 0] = 0.0
    return yt.load_amr_grids(grids, dims)


@pytest.mark.parametrize("level,start,dims", [(0, (-1, -1, -1), (6, 6, 6)), (0, (3, 5, -1), (6, 5, 7)), (1, (3, 3, 3), (6, 6, 6)),
                                              (2, (11, 12, 13), (6, 6, 6))])
def test_fill_region_equals_the_reference(level, start, dims):
    _, fr = _ref_lib()
    ds = _nested()
    key = ds._resolve("density")
    got = loaders.fill_region(ds, key, level, start, dims)
    # what a data source with min_level = max_level = level hands to fill_region: every cell of that level's grids
    ipos, vals = [], []
    for g in ds.index.grids:
        if g.Level == level:
            idx = np.indices(tuple(g.ActiveDimensions)).reshape(3, -1).T + g.get_global_startindex()[None, :]
            ipos.append(idx)
            vals.append(g.fields[key].ravel())
    ipos = np.ascontiguousarray(np.concatenate(ipos), dtype="int64")
    vals = np.ascontiguousarray(np.concatenate(vals), dtype="float64")
    want = np.zeros(tuple(dims), dtype="float64")
    level_dims = (ds.domain_dimensions * ds.refine_by ** level).astype("int64")
    fr.fill_region([vals], [want], level, np.array(start, "int64"), ipos, np.full(len(vals), level, "int64"), level_dims,
                   np.array([ds.refine_by] * 3, "int64"))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("nprocs", [1, 4])
def test_one_level_is_the_periodic_global_array(nprocs):
    arr = np.random.default_rng(1).random((8, 6, 10)) + 0.1
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=nprocs)
    want = loaders.vertex_centre_from_ghost(np.pad(arr, 1, mode="wrap"))
    for g in ds.index.grids:
        v = g.get_vertex_centered_data([("stream", "density")], no_ghost=False)[("stream", "density")]
        lo = g.get_global_startindex()
        hi = lo + g.ActiveDimensions
        assert np.array_equal(v, want[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1])


def test_non_periodic_ghost_zones_beyond_the_domain_are_zero():
    arr = np.random.default_rng(1).random((6, 6, 6)) + 0.1
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, periodicity=(False, True, False))
    cg = ds.index.grids[0].retrieve_ghost_zones(1, ["density"], smoothed=True)[("stream", "density")]
    assert np.all(cg[0] == 0) and np.all(cg[-1] == 0) and np.all(cg[:, :, 0] == 0)
    assert np.array_equal(cg[1:-1, 0, 1:-1], arr[:, -1, :])


def test_nested_levels_same_level_neighbours_and_coarse_interpolation():
    ds = _nested()
    key = ds._resolve("density")
    fine = [g for g in ds.index.grids if g.Level == 2]
    for g in fine[:6] + fine[-3:]:
        cg = g.retrieve_ghost_zones(1, ["density"], smoothed=True)[key]
        assert cg.shape == tuple(g.ActiveDimensions + 2)
        assert np.array_equal(cg[1:-1, 1:-1, 1:-1], g.fields[key])           # the grid's own cells
        start = g.get_global_startindex() - 1
        same = loaders.fill_region(ds, key, 2, start, g.ActiveDimensions + 2)    # cells of level-2 grids
        covered = same != 0
        assert np.array_equal(cg[covered], same[covered])
        # the rest comes from the level-1 box by ghost_zone_interpolate, itself level-1 cells over interpolated level 0
        assert np.all(np.isfinite(cg)) and np.all(cg[~covered] > 0)
        lo, hi = ds.index.grids[0].fields[key].min(), max(gg.fields[key].max() for gg in ds.index.grids)
        assert cg.min() >= 0 and cg.max() <= hi * 1.5 + lo
    # a vertex array has one more entry per axis and averages the padded cells
    v = fine[0].get_vertex_centered_data(["density"], no_ghost=False)[key]
    cg = fine[0].retrieve_ghost_zones(1, ["density"], smoothed=True)[key]
    assert v.shape == tuple(fine[0].ActiveDimensions + 1)
    assert np.allclose(v[2, 1, 3], cg[2:4, 1:3, 3:5].mean(), rtol=1e-14)


def test_zero_cells_do_not_replace_interpolated_values():
    ds = _nested(zero_cells=True)
    key = ds._resolve("density")
    g = ds.index.grids[-1]
    cg = g.retrieve_ghost_zones(1, ["density"], smoothed=True)[key]
    inner = cg[1:-1, 1:-1, 1:-1]
    assert np.all(g.fields[key][1:3, 2, 0] == 0.0)
    assert np.all(inner[1:3, 2, 0] > 0.0)                 # interpolated from level 1 instead (output != 0 test)
    m = g.fields[key] != 0
    assert np.array_equal(inner[m], g.fields[key][m])


@pytest.mark.gpu
@pytest.mark.parametrize("amr", [False, True])
def test_device_ghost_bricks_equal_the_host_path_bit_for_bit(amr, gpu):
    import torch
    from yt_b200 import _lib as L
    from yt_b200.amr_kdtree import AMRKDTree
    if amr:
        ds = _nested()
        log = [True]
    else:
        arr = np.random.default_rng(5).random((12, 10, 14)) + 0.2
        ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=4)
        log = [False]
    kd = AMRKDTree(ds)
    kd.set_fields([ds.default_field], log, False)          # no_ghost = False
    bricks = kd.device_bricks(gpu, L.STORE_F64)
    leaves = [int(i) for i in kd.tree.traverse()]
    assert len(leaves) > 1
    from yt_b200.parallel import _DevArray
    for slot, node in enumerate(leaves):
        host = kd.get_brick_data(node).my_data[0]
        ptrs, _ = bricks.device_ptrs(slot)
        px, py = bricks.pitch(slot)
        shape = host.shape
        dev = torch.as_tensor(_DevArray(ptrs[0], (shape[0], px // py, py)), device="cuda:%d" % gpu.device)
        got = dev[:, :shape[1], :shape[2]].cpu().numpy()
        if amr:   # log10: CUDA's is within 2 ulp of libm's
            assert np.allclose(got, host, rtol=4e-16, atol=0, equal_nan=True)
        else:
            assert np.array_equal(got, host)


@pytest.mark.gpu
def test_scene_with_ghost_zones_matches_the_oracle(gpu):
    from tests import common as T
    from tests.test_scene_gpu import oracle_render
    arr = np.random.default_rng(11).random((24, 24, 24)) + 0.3
    ds = yt.load_uniform_grid({"density": arr}, arr.shape, nprocs=8)
    sc = yt.create_scene(ds)
    sc.camera.resolution = (96, 80)
    src = sc.get_source()
    src.use_ghost_zones = True
    im = np.asarray(sc.render())
    want, nb = oracle_render(sc)
    assert nb == 8 and want[:, :, :3].max() > 0
    assert T.normalised_diff(im, want, False) <= 1e-4
    # and it is not the no-ghost image: the vertex values on internal faces differ
    src2 = yt.create_scene(ds)
    src2.camera.resolution = (96, 80)
    im2 = np.asarray(src2.render())
    assert np.abs(im2 - im).max() > 0
