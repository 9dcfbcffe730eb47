"""Vertex centring: the oracle's restatement of ghost_zone_interpolate against the
compiled reference and the golden vectors (CPU), and the CUDA kernel behind ytb200_brick_from_cells against
the oracle (GPU).  fp64 storage of a linear field must be BIT-identical; log10 is CUDA's (<= 2 ulp from libm)."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden.make_vertex_golden import SHAPES, cells_for

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vertex_golden.npz")


@pytest.mark.parametrize("k", range(len(SHAPES)))
def test_oracle_matches_golden_bit_for_bit(k):
    gold = np.load(GOLD)
    got = O.vertex_centre(cells_for(SHAPES[k], k))
    assert np.array_equal(got, gold["v%d" % k])


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape", [(4, 4, 4), (7, 3, 5), (2, 9, 2), (12, 12, 12)])
def test_oracle_matches_compiled_reference_bit_for_bit(shape):
    ref_dir = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from yt.utilities.lib.interpolators import ghost_zone_interpolate
    cells = np.random.default_rng(3).normal(size=shape) * 1e3
    want = np.zeros(tuple(s + 1 for s in shape))
    ghost_zone_interpolate(1, cells, np.array([0.5] * 3), want, np.zeros(3))
    assert np.array_equal(O.vertex_centre(cells), want)


def test_brick_data_log_and_slice():
    cells = cells_for((6, 5, 4), 7)
    full = O.vertex_centre(cells)
    got = O.brick_data_from_cells(cells, (1, 0, 2), (4, 5, 4), True)
    assert got.shape == (4, 6, 3)
    sub = full[1:5, 0:6, 2:5]
    assert np.array_equal(np.isnan(got), sub <= 0)
    ok = sub > 0
    assert np.array_equal(got[ok], np.log10(sub[ok]))


# ---------------------------------------------------------------------------------------------------
def _device_brick(cells, lo, hi, take_log, storage, cells_f32=False):
    import torch
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200.parallel import _DevArray
    ctx = D.get_context()
    bs = D.BrickSet(ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F64 if storage == "f64" else L.STORE_F32)
    if cells_f32:
        t = torch.as_tensor(cells.astype("float32"), device="cuda:%d" % ctx.device).contiguous()
        ptr, keep = t.data_ptr(), t
    else:
        dc = D.DeviceCells(ctx, cells)
        ptr, keep = dc.ptr, dc
    n = np.array(cells.shape)
    bs.from_cells(0, [ptr], n, lo, hi, np.array(lo) / n, np.array(hi) / n, log_fields=[take_log], cells_are_f32=cells_f32, keep=keep)
    bs.commit()
    ctx.synchronize()
    shape = tuple(int(h - l + 1) for l, h in zip(lo, hi))
    return read_brick(bs, 0, shape, storage, ctx)


def read_brick(bs, index, shape, storage, ctx):
    """Device array of a brick as a contiguous numpy array (library-owned arrays may have padded rows)."""
    import torch
    from yt_b200.parallel import _DevArray
    ptrs, _ = bs.device_ptrs(index)
    _, row = bs.pitch(index)
    arr = torch.as_tensor(_DevArray(ptrs[0], (shape[0], shape[1], row), "<f8" if storage == "f64" else "<f4"), device="cuda:%d" % ctx.device)
    return arr[:, :, :shape[2]].cpu().numpy().copy()


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(SHAPES)))
def test_device_vertex_centring_fp64_is_bit_identical(gpu, k):
    cells = cells_for(SHAPES[k], k)
    want = O.vertex_centre(cells)
    got = _device_brick(cells, (0, 0, 0), cells.shape, False, "f64")
    assert np.array_equal(got, want)
    assert np.array_equal(got, np.load(GOLD)["v%d" % k])


@pytest.mark.gpu
def test_device_vertex_centring_slice_log_and_fp32(gpu):
    cells = cells_for((16, 12, 20), 11)
    lo, hi = (3, 0, 7), (11, 12, 20)
    want = O.brick_data_from_cells(cells, lo, hi, True)
    got = _device_brick(cells, lo, hi, True, "f64")
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    # CUDA log10 vs libm log10: both within 1 ulp of the true value
    assert np.all(np.abs(got[ok] - want[ok]) <= 4.5e-16 * np.maximum(np.abs(want[ok]), 1e-300) + 1e-300)
    got32 = _device_brick(cells, lo, hi, True, "f32")
    assert got32.dtype == np.float32
    assert np.array_equal(np.isnan(got32), np.isnan(want))
    assert np.all(np.abs(got32[ok].astype("float64") - want[ok]) <= 6.0e-8 * np.abs(want[ok]) + 1e-30)
    # linear field, fp32 storage: exactly the rounded fp64 result
    lin = _device_brick(cells, lo, hi, False, "f32")
    assert np.array_equal(lin, O.brick_data_from_cells(cells, lo, hi, False).astype("float32"))


@pytest.mark.gpu
def test_device_vertex_centring_from_fp32_cells(gpu):
    cells = cells_for((9, 10, 8), 12).astype("float32").astype("float64")   # exactly representable in fp32
    want = O.vertex_centre(cells)
    got = _device_brick(cells, (0, 0, 0), cells.shape, False, "f64", cells_f32=True)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["f32", "f64"])
def test_rows_of_a_multiple_of_256_bytes_are_padded(gpu, storage):
    """63 cells -> 64 vertices per row = 256 B (fp32) / 512 B (fp64): the library pads the row by one sector and the
    values are unchanged."""
    cells = cells_for((5, 6, 63), 13)
    want = O.vertex_centre(cells)
    got = _device_brick(cells, (0, 0, 0), cells.shape, False, storage)
    assert np.array_equal(got, want.astype("float32") if storage == "f32" else want)


@pytest.mark.gpu
def test_one_cell_axis_is_refused(gpu):
    from yt_b200 import _lib as L
    with pytest.raises(L.YTB200Error, match="2 cells"):
        _device_brick(np.ones((1, 4, 4)), (0, 0, 0), (1, 4, 4), False, "f64")


@pytest.mark.gpu
def test_kdtree_device_route_equals_host_route(gpu):
    """AMRKDTree.device_bricks: bricks centred on the device hold the values get_brick_data builds on the host
    (log field, AMR hierarchy, fp64 storage: equal up to the summation order of the host mirror / log10 ulps)."""
    import torch
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import loaders, synthetic
    from yt_b200.amr_kdtree import AMRKDTree
    from yt_b200.parallel import _DevArray
    grids, dd = synthetic.nested_amr_grids(base_cells=16, tile=8, levels=3, seed=2)
    ds = loaders.load_amr_grids(grids, dd)
    kd = AMRKDTree(ds)
    kd.set_fields([ds.default_field], [True], True)
    ctx = D.get_context()
    dev = kd.device_bricks(ctx, L.STORE_F64, centre_on_device=True)
    leaves = [int(i) for i in kd.tree.traverse()]
    assert dev.n == len(leaves) > 8
    for slot in (0, 1, len(leaves) // 2, len(leaves) - 1):
        host = kd.get_brick_data(leaves[slot]).my_data[0]
        got = read_brick(dev, slot, host.shape, "f64", ctx)
        assert np.array_equal(np.isnan(got), np.isnan(host))
        ok = ~np.isnan(host)
        assert np.allclose(got[ok], host[ok], rtol=1e-14, atol=1e-14)
