"""Parity of the CUDA samplers (through the C ABI) against the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star): ProjectionSampler / InterpolatedProjectionSampler integrals within
rtol 1e-6; VolumeRenderSampler images within 1e-4 max-abs on the normalised RGBA.
EXACT mode (fp64, reference evaluation order) must agree to rounding noise.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_images.npz")
PROJ_RTOL = 1e-6      # north_star tolerance for projection integrals
VR_ATOL = 1e-4        # north_star tolerance for normalised volume-rendered RGBA


def _gpu_mods():
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    return S, G


def _check(sc, got, want, exact):
    assert got.shape == want.shape
    assert not np.isnan(got).any()
    if sc.kind == "vr":
        err = T.normalised_diff(got, want, sc.grey)
        assert err <= (1e-10 if exact else VR_ATOL), (sc.name, err)
        if not sc.grey:  # alpha untouched by the sampler (H9)
            assert np.array_equal(got[..., 3], want[..., 3])
    else:
        scale = np.abs(want).max()
        err = np.abs(got - want).max() / scale
        assert err <= (1e-13 if exact else PROJ_RTOL), (sc.name, err)
        # per-pixel relative check where the integral is not tiny
        big = np.abs(want) > 1e-3 * scale
        rel = np.abs(got[big] - want[big]) / np.abs(want[big])
        assert rel.max() <= (1e-12 if exact else PROJ_RTOL), (sc.name, rel.max())


@pytest.mark.parametrize("sc", T.scenarios(), ids=lambda s: s.name)
def test_fast_path_matches_oracle(sc, gpu):
    S, G = _gpu_mods()
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G, sampler_kwargs={"storage": "f32", "precision": "fast"})
    _check(sc, got, want, exact=False)


@pytest.mark.parametrize("sc", T.scenarios(), ids=lambda s: s.name)
def test_exact_path_matches_oracle_to_rounding(sc, gpu):
    S, G = _gpu_mods()
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G, sampler_kwargs={"storage": "f64", "precision": "exact"})
    _check(sc, got, want, exact=True)


@pytest.mark.parametrize("sc", T.scenarios(), ids=lambda s: s.name)
def test_fast_path_matches_golden(sc, gpu):
    S, G = _gpu_mods()
    want = np.load(GOLD)[sc.name]
    got = T.run_scenario(sc, S, G)
    _check(sc, got, want, exact=False)


@pytest.mark.parametrize("sc", [s for s in T.scenarios() if s.kind == "vr"], ids=lambda s: s.name)
def test_fast_path_without_the_line_form_tables(sc, gpu, monkeypatch):
    """Tables that pass the accuracy guard of ytb200_tf_create sample through LINE-form rows (y = d * fs + c); a table that
    fails it keeps (value, increment) rows.  YTB200_NO_LINE forces the latter for every table: both forms must give the
    oracle's image, and they agree with each other far inside the bar."""
    S, G = _gpu_mods()
    want = T.run_scenario(sc, O, O, num_threads=0)
    with_line = T.run_scenario(sc, S, G)
    monkeypatch.setenv("YTB200_NO_LINE", "1")
    without = T.run_scenario(sc, S, G)
    _check(sc, without, want, exact=False)
    _check(sc, with_line, want, exact=False)
    assert T.normalised_diff(with_line, without, sc.grey) <= 2e-5


@pytest.mark.parametrize("sc", [s for s in T.scenarios() if s.kind != "proj"][:6], ids=lambda s: s.name)
def test_f64_storage_fast_arithmetic(sc, gpu):
    S, G = _gpu_mods()
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G, sampler_kwargs={"storage": "f64", "precision": "fast"})
    _check(sc, got, want, exact=False)


def test_sample_count_matches_oracle(gpu):
    """The counting kernel returns exactly the number of cells the reference's walk_volume visits."""
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    for sc in T.scenarios():
        if sc.name not in ("proj_corner", "vr_mask_nan", "vr_8bricks_pp", "proj_on_axis_edges"):
            continue
        img = np.zeros((sc.cam.resolution[0], sc.cam.resolution[1], 4))
        args, kwargs = T.sampler_args(sc.cam, image=img)
        kwargs.pop("camera_data")
        if sc.kind == "vr":
            o = O.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs)
            s = S.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs)
        else:
            o = O.ProjectionSampler(*args, sc.n_samples, **kwargs)
            s = S.ProjectionSampler(*args, sc.n_samples, **kwargs)
        og = [O.PartitionedGrid(i, list(b[0]), b[1], b[2], b[3], np.array(b[4], "int64")) for i, b in enumerate(sc.bricks)]
        gg = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
              for i, b in enumerate(sc.bricks)]
        for i in sc.order:
            o(og[i], num_threads=1)
        bs = G.build_brickset(gg)
        cells, samples = s.count_samples(bs, sc.order)
        assert cells == o.cells, (sc.name, cells, o.cells)
        assert samples == cells * (sc.n_samples if sc.kind != "proj" else 1)


def test_flush_is_incremental(gpu):
    """Reading .aimage between bricks and continuing gives the same image as one batch."""
    S, G = _gpu_mods()
    sc = [s for s in T.scenarios() if s.name == "vr_8bricks_perspective_grey"][0]
    one = T.run_scenario(sc, S, G)
    img = np.zeros_like(one)
    args, kwargs = T.sampler_args(sc.cam, image=img)
    s = S.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs)
    grids = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
             for i, b in enumerate(sc.bricks)]
    for k, i in enumerate(sc.order):
        s(grids[i])
        if k % 3 == 0:
            assert s.aimage is img
    assert np.allclose(s.aimage, one, rtol=0, atol=2e-6)


def test_reference_error_behaviour(gpu):
    S, G = _gpu_mods()
    sc = T.scenarios()[0]
    args, kwargs = T.sampler_args(sc.cam)
    kwargs.pop("camera_data")
    ro = args[0].copy()
    ro.setflags(write=False)
    with pytest.raises(ValueError, match="read-only"):
        S.ProjectionSampler(ro, *args[1:], 10, **kwargs)
    with pytest.raises(ValueError):
        S.ProjectionSampler(*args[:4], args[4].astype("float32"), *args[5:], 10, **kwargs)
    s = S.ProjectionSampler(*args[:8], "Nonsense", 10, **kwargs)
    b = sc.bricks[0]
    pg = G.PartitionedGrid(0, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
    with pytest.raises(NotImplementedError):
        s(pg)
    persp = [x for x in T.scenarios() if x.name == "proj_perspective"][0]
    pargs, pkw = T.sampler_args(persp.cam)
    with pytest.raises(RuntimeError, match="Bad lens shape"):
        S.ProjectionSampler(pargs[0][:-1], *pargs[1:], 10, **pkw)


@pytest.mark.parametrize("name", ["vr_colour_grey_zbuf", "proj_offaxis_zbuf", "vr_8bricks_perspective_grey", "vr_colour"])
def test_render_in_place_on_pinned_host_arrays(gpu, name, monkeypatch):
    """ytb200_render with page-locked host arrays (the kernel reads/writes them over PCIe) gives the same
    image as the staged route (copy in, cast, copy out) and matches the oracle."""
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200.partitioned_grid import build_brickset
    S, G = _gpu_mods()
    sc = [s for s in T.scenarios() if s.name == name][0]
    want = T.run_scenario(sc, O, O, num_threads=0)
    nx, ny = sc.cam.resolution
    outs = []
    for zero_copy in (True, False):
        if zero_copy:
            monkeypatch.delenv("YTB200_NO_ZEROCOPY", raising=False)
        else:
            monkeypatch.setenv("YTB200_NO_ZEROCOPY", "1")
        img = D.pinned_empty((nx, ny, 4))
        img[:] = 0.0 if sc.image0 is None else sc.image0
        zb = None
        if sc.zbuffer is not None:
            zb = D.pinned_empty((nx, ny))
            zb[:] = sc.zbuffer
        args, kwargs = T.sampler_args(sc.cam, zbuffer=zb, image=img)
        if sc.cam.lens.lens_type != "perspective" or sc.kind != "vr":
            kwargs.pop("camera_data")
        if sc.kind == "vr":
            s = S.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs)
        else:
            s = S.ProjectionSampler(*args, sc.n_samples, **kwargs)
        grids = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
                 for i, b in enumerate(sc.bricks)]
        bs = build_brickset(grids, s._ctx, s._storage)
        s._frame.render(s.kind, bs, np.array(sc.order), img, zbuffer=zb, tf=s._tables, n_samples=sc.n_samples)
        outs.append(np.array(img, copy=True))
    assert np.array_equal(outs[0], outs[1])
    _check(sc, outs[0], want, exact=False)
