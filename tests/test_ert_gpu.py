"""Early ray termination (ours; ``ytb200_frame_set_early_termination``): the grey-opacity ColorTransferFunction cast that
evaluates FIT_eval_transfer's composition (field_interpolation_tables.pxd:103-110) from its far end and stops a ray once
what it has composited transmits less than eps.  It must give the reference's image (1e-4 on the normalised RGBA, like
every volume render), the image of the plain cast to within eps, and it must actually skip work on opaque tables."""
import copy

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

pytestmark = pytest.mark.gpu
EPS = 1e-6


def _cases():
    base = {s.name: s for s in T.scenarios()}
    out = []
    rng = np.random.RandomState(123)
    # one brick with initial colours and a z-buffer (max_t < 1: the reversed ray starts at the opaque surface),
    # eight bricks in view order through a perspective lens; translucent, opaque, and opaque enough to clamp
    for name, height in (("vr_colour_grey_zbuf", 40.0), ("vr_colour_grey_zbuf", 400.0), ("vr_colour_grey_zbuf", 30000.0),
                         ("vr_8bricks_perspective_grey", 60.0), ("vr_8bricks_perspective_grey", 2000.0)):
        sc = copy.copy(base[name])
        sc.tf = T.colour_tf(grey=True, height=height)
        if sc.image0 is None:
            nx, ny = sc.cam.resolution
            sc.image0 = rng.uniform(0, 0.4, size=(nx, ny, 4))
        sc.name = "%s_h%g" % (name, height)
        out.append(sc)
    return out


def _sampler(sc, ert):
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    img = sc.image0.copy() if sc.image0 is not None else np.zeros(tuple(sc.cam.resolution) + (4,))
    args, kwargs = T.sampler_args(sc.cam, zbuffer=None if sc.zbuffer is None else sc.zbuffer.copy(), image=img)
    if sc.cam.lens.lens_type != "perspective":
        kwargs.pop("camera_data")
    s = S.VolumeRenderSampler(*args, sc.tf, sc.n_samples, early_termination=ert, **kwargs)
    grids = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
             for i, b in enumerate(sc.bricks)]
    return s, G.build_brickset(grids, s._ctx, s._storage), img


@pytest.mark.parametrize("sc", _cases(), ids=lambda s: s.name)
def test_terminated_cast_gives_the_reference_image(sc, gpu):
    want = T.run_scenario(sc, O, O, num_threads=0)
    order = np.array(sc.order, dtype=np.int64)
    images, cells = {}, {}
    for ert in (0.0, EPS):
        s, bs, img = _sampler(sc, ert)
        fr = s._frame
        fr.upload(img, s.azbuffer)
        st = fr.cast(s.kind, bs, order, tf=s._tables, n_samples=s.n_samples, timed=True)
        images[ert] = fr.download(np.empty_like(img))
        cells[ert] = int(st.n_cells)
        total = int(fr.count(s.kind, bs, order, s.n_samples).n_cells)
    err = T.normalised_diff(images[EPS], want, True)
    assert err <= 1e-4, (sc.name, err)
    # against the plain cast: the same samples composited from the other end, minus what hides behind B < eps
    scale = max(np.abs(images[0.0]).max(), 1.0)
    assert np.abs(images[EPS] - images[0.0]).max() <= 2e-5 * scale, sc.name
    assert 0 < cells[EPS] <= total
    print("%s: %d of %d cells sampled (%.0f %% skipped), normalised diff to the oracle %.2g" %
          (sc.name, cells[EPS], total, 100.0 * (1.0 - cells[EPS] / total), err))
    if "h30000" in sc.name or "h2000" in sc.name:
        assert cells[EPS] < 0.6 * total, (cells[EPS], total)     # an opaque table must actually stop rays


def test_per_brick_calls_and_whole_list_agree(gpu):
    """The reference API hands the sampler one brick at a time; every call then terminates inside its own brick."""
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    sc = _cases()[4]
    want = T.run_scenario(sc, O, O, num_threads=0)
    got = T.run_scenario(sc, S, G, sampler_kwargs={"early_termination": EPS})
    assert T.normalised_diff(got, want, True) <= 1e-4


def test_setting_is_ignored_where_it_cannot_apply(gpu):
    """A non-grey table absorbs each channel with its own emission: the plain cast runs, bit for bit."""
    base = {s.name: s for s in T.scenarios()}
    sc = base["vr_8bricks_pp"]
    order = np.array(sc.order, dtype=np.int64)
    out = []
    for ert in (0.0, EPS):
        s, bs, img = _sampler(sc, ert)
        s._frame.upload(img, s.azbuffer)
        s._frame.cast(s.kind, bs, order, tf=s._tables, n_samples=s.n_samples)
        out.append(s._frame.download(np.empty_like(img)))
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("height", [60.0, 30000.0])
def test_terminated_partials_compose_like_the_plain_cast(height, gpu):
    """The multi-GPU shape of it: two halves of the list cast as affine partials with early termination, composed and
    applied to the initial colours."""
    import torch
    from yt_b200 import _lib as L
    base = {s.name: s for s in T.scenarios()}
    sc = copy.copy(base["vr_8bricks_perspective_grey"])
    sc.tf = T.colour_tf(grey=True, height=height)
    nx, ny = sc.cam.resolution
    sc.image0 = np.random.RandomState(5).uniform(0, 0.4, size=(nx, ny, 4))
    want = T.run_scenario(sc, O, O, num_threads=0)
    s, bs, img = _sampler(sc, EPS)
    fr, ctx = s._frame, s._ctx
    fr.upload(img, s.azbuffer)
    npx = nx * ny
    order = np.array(sc.order, dtype=np.int64)
    parts = [order[:4], order[4:]]
    bufs = [torch.zeros((npx, 8), dtype=torch.float64, device="cuda:%d" % ctx.device) for _ in parts]
    torch.cuda.synchronize()
    for o, b in zip(parts, bufs):
        fr.cast_partial(s.kind, bs, o, b.data_ptr(), tf=s._tables, n_samples=s.n_samples)
    lib = L.load()
    L.check(lib.ytb200_partial_compose(ctx.handle, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[0].data_ptr(), npx))
    L.check(lib.ytb200_partial_apply(ctx.handle, bufs[0].data_ptr(), fr.device_image, npx))
    got = fr.download(np.empty_like(img))
    assert T.normalised_diff(got, want, True) <= 1e-4


@pytest.mark.parametrize("height", [60.0, 30000.0])
def test_terminated_cast_through_the_swap_object(height, gpu):
    """The compact partial format of the grey render (fp32 A only, B = 1 - A_alpha or 0 where a clamp fired) must carry
    the terminated cast's map too: there B is tracked explicitly and A_alpha = 1 - B by construction."""
    from yt_b200 import _lib as L
    from yt_b200.parallel import PeerSwap, partial_format
    base = {s.name: s for s in T.scenarios()}
    sc = copy.copy(base["vr_8bricks_perspective_grey"])
    sc.tf = T.colour_tf(grey=True, height=height)
    nx, ny = sc.cam.resolution
    sc.image0 = np.random.RandomState(6).uniform(0, 0.4, size=(nx, ny, 4))
    want = T.run_scenario(sc, O, O, num_threads=0)
    s, bs, img = _sampler(sc, EPS)
    fr = s._frame
    fmt = partial_format(s.kind, s._tables, bs.n_fields)
    assert fmt == L.PARTIAL_A32G
    swap = PeerSwap(s._ctx, fr, [], sc.cam.lens.viewpoint, 0, 1, fmt=fmt)
    for _ in range(2):
        fr.upload(img, s.azbuffer)
        swap.cast(fr, s.kind, bs, np.array(sc.order, dtype=np.int64), tf=s._tables, n_samples=s.n_samples)
        swap.composite(fr)
        swap.check()
        got = fr.download(np.empty_like(img))
        assert T.normalised_diff(got, want, True) <= 1e-4


@pytest.mark.parametrize("height", [300.0, 30000.0])
def test_terminated_cast_over_a_long_brick_list_with_masked_and_nan_cells(height, gpu):
    """More than 32 bricks select the long-list instantiation of the cast (the one AMR hierarchies use); here 64 bricks of
    4^3 cells in view order through a perspective lens, with masked cells and NaN / zero vertices in the data."""
    base = {s.name: s for s in T.scenarios()}
    src = base["vr_mask_nan"]                      # data with NaN and 0 vertices, a mask with holes
    data, mask, le, re, dims = src.bricks[0]
    sub = T.split_brick(data, mask, le, re, (4, 4, 4))
    cam = T.make_camera("perspective", resolution=(48, 48), position=[1.6, 1.4, 1.2], width=[0.6, 0.6, 1.0])
    order = T.view_order([(b[2], b[3]) for b in sub], cam.lens.viewpoint)
    assert len(sub) == 64
    img0 = np.random.RandomState(9).uniform(0, 0.3, size=(48, 48, 4))
    sc = T.Scenario("ert_64bricks", "vr", cam, sub, order, tf=T.colour_tf(grey=True, height=height), image0=img0, grey=True)
    want = T.run_scenario(sc, O, O, num_threads=0)
    cells = {}
    for ert in (0.0, EPS):
        s, bs, img = _sampler(sc, ert)
        fr = s._frame
        fr.upload(img, s.azbuffer)
        st = fr.cast(s.kind, bs, np.array(sc.order, dtype=np.int64), tf=s._tables, n_samples=s.n_samples, timed=True)
        got = fr.download(np.empty_like(img))
        cells[ert] = int(st.n_cells)
        ok = ~np.isnan(want)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert T.normalised_diff(np.where(ok, got, 0.0), np.where(ok, want, 0.0), True) <= 1e-4, (height, ert)
    if height > 1000:
        assert 0 < cells[EPS]
