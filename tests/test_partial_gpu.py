"""Affine partial images on ONE GPU: what every rank of a multi-GPU render computes.
Casting the first and the second half of a view-ordered brick list into two partials C -> A + B*C, composing them and
applying the result to the initial colours must give the image of the plain cast -- including when opacity clamps
fire (the colour path derives B from A and a sticky per-channel flag) and with fp32 partials through the swap object."""
import copy

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as T

pytestmark = pytest.mark.gpu


def _cases():
    base = {s.name: s for s in T.scenarios()}
    out = []
    rng = np.random.RandomState(77)
    for name, height in (("vr_8bricks_pp", 100.0), ("vr_8bricks_pp", 40000.0),
                         ("vr_8bricks_perspective_grey", 60.0), ("vr_8bricks_perspective_grey", 30000.0)):
        sc = copy.copy(base[name])
        nx, ny = sc.cam.resolution
        sc.tf = T.colour_tf(grey=sc.grey, height=height)
        sc.image0 = rng.uniform(0, 0.4, size=(nx, ny, 4))
        sc.zbuffer = rng.uniform(0.5, 1.3, size=(nx, ny))
        sc.name = "%s_h%g" % (name, height)
        out.append(sc)
    sc = copy.copy(base["proj_corner"])
    sc.name = "proj_two_fields"
    out.append(sc)
    return out


def _setup(sc):
    from yt_b200 import image_samplers as S
    from yt_b200 import partitioned_grid as G
    img = sc.image0.copy() if sc.image0 is not None else np.zeros(tuple(sc.cam.resolution) + (4,))
    args, kwargs = T.sampler_args(sc.cam, zbuffer=None if sc.zbuffer is None else sc.zbuffer.copy(), image=img)
    if sc.cam.lens.lens_type != "perspective" or sc.kind != "vr":
        kwargs.pop("camera_data")
    s = S.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs) if sc.kind == "vr" else S.ProjectionSampler(*args, sc.n_samples, **kwargs)
    grids = [G.PartitionedGrid(i, list(b[0]), b[1], np.array(b[2], "float64"), np.array(b[3], "float64"), np.array(b[4], "int64"))
             for i, b in enumerate(sc.bricks)]
    bs = G.build_brickset(grids, s._ctx, s._storage)
    return s, bs, img


def _check(sc, got, want):
    if sc.kind == "vr":
        err = T.normalised_diff(got, want, sc.grey)
        assert err <= 1e-4, (sc.name, err)
        if not sc.grey:
            assert np.array_equal(got[..., 3], want[..., 3])   # non-grey never touches alpha (H9)
    else:
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), sc.name


@pytest.mark.parametrize("sc", _cases(), ids=lambda s: s.name)
def test_two_partials_composed_equal_the_plain_cast(sc, gpu):
    import torch
    from yt_b200 import _lib as L
    want = T.run_scenario(sc, O, O, num_threads=0)
    s, bs, img = _setup(sc)
    fr, ctx = s._frame, s._ctx
    fr.upload(img, s.azbuffer)
    npx = img.shape[0] * img.shape[1]
    dev = "cuda:%d" % ctx.device
    order = np.array(sc.order, dtype=np.int64)
    half = max(len(order) // 2, 1) if len(order) > 1 else 1
    parts = [order[:half], order[half:]]
    bufs = [torch.zeros((npx, 8), dtype=torch.float64, device=dev) for _ in parts]
    torch.cuda.synchronize()
    for o, b in zip(parts, bufs):
        fr.cast_partial(s.kind, bs, o, b.data_ptr(), tf=s._tables, n_samples=s.n_samples)
    lib = L.load()
    L.check(lib.ytb200_partial_compose(ctx.handle, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[0].data_ptr(), npx))
    L.check(lib.ytb200_partial_apply(ctx.handle, bufs[0].data_ptr(), fr.device_image, npx))
    got = fr.download(np.empty_like(img))
    _check(sc, got, want)
    if sc.kind == "vr" and "h4" in sc.name or "h3" in sc.name:
        # the opaque cases really do clamp: some transmittance is exactly zero
        B = bufs[0].cpu().numpy()[:, 4:7]
        assert (B == 0.0).any()


@pytest.mark.parametrize("sc", _cases(), ids=lambda s: s.name)
@pytest.mark.parametrize("compact", [True, False])
def test_partial_through_the_swap_object(sc, compact, gpu):
    """world = 1: no rounds, the swap's apply kernel decodes the partial (compact: 16 B per pixel for the colour render,
    8 B per summed field for the projection; else the general fp32 / fp64 (A, B) pair) and applies it to the initial
    colours -- including pixels whose opacity clamps fired (their flag must survive the encoding)."""
    from yt_b200 import _lib as L
    from yt_b200.parallel import PeerSwap, partial_format
    want = T.run_scenario(sc, O, O, num_threads=0)
    s, bs, img = _setup(sc)
    fr = s._frame
    fmt = partial_format(s.kind, s._tables, bs.n_fields)
    if sc.kind == "vr":
        assert fmt == (L.PARTIAL_A32G if sc.grey else L.PARTIAL_A32)
    else:
        assert fmt == L.PARTIAL_SUM2
    if not compact:
        fmt = L.PARTIAL_AB32 if sc.kind == "vr" else L.PARTIAL_SUM4
    swap = PeerSwap(s._ctx, fr, [], sc.cam.lens.viewpoint, 0, 1, fmt=fmt)
    for _ in range(3):   # three frames: both buffer parities are used
        fr.upload(img, s.azbuffer)
        swap.cast(fr, s.kind, bs, np.array(sc.order, dtype=np.int64), tf=s._tables, n_samples=s.n_samples)
        swap.composite(fr)
        swap.check()
        got = fr.download(np.empty_like(img))
        _check(sc, got, want)


def test_swap_composite_into_a_page_locked_host_image(gpu):
    """composite_host: the finished pixels go straight into a registered host array holding the initial colours."""
    from yt_b200.parallel import PeerSwap, partial_format, shared_pinned_empty
    sc = _cases()[2]
    want = T.run_scenario(sc, O, O, num_threads=0)
    s, bs, img = _setup(sc)
    fr = s._frame
    host = shared_pinned_empty(img.shape)
    swap = PeerSwap(s._ctx, fr, [], sc.cam.lens.viewpoint, 0, 1, fmt=partial_format(s.kind, s._tables, bs.n_fields))
    for _ in range(2):
        host[:] = img
        fr.upload(img, s.azbuffer)
        swap.cast(fr, s.kind, bs, np.array(sc.order, dtype=np.int64), tf=s._tables, n_samples=s.n_samples)
        swap.composite_host(fr, host)
        swap.check()
        _check(sc, np.array(host), want)
