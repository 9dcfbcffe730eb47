"""Oracle differentials AT BASELINE.json's sizes: the frames the benchmark times, against the reference samplers.

The CUDA cast renders the full frame; the compiled reference (``oracle/_ref``, the unmodified Cython samplers -- it
travels to the GPU box; the C oracle, pinned bitwise to it, stands in if it is absent) casts a LATTICE of the very same
rays through the very same fp32-rounded data widened to fp64:

* the reference takes any ray set through its per-pixel ``vp_pos`` / ``vp_dir`` arrays (``generate_vector_info_null``,
  lenses.pyx:196-207), so a (64, 64) sub-image holding every 32nd pixel's plane-parallel ray
  (``generate_vector_info_plane_parallel``, lenses.pyx:179-192, same operation order) costs ~7e7 samples instead of 3e10;
* the 1024^3 volume is handed to it as x-slabs in increasing ray parameter (what ``walk_volume`` does inside one brick,
  SURVEY.md finding 6 / H14), downloaded one at a time, so the host never holds more than ~1.1 GB of fp64.

Bars: volume render max-abs 1e-4 on the normalised RGBA, projection rtol 1e-6 of the largest integral.
"""
import numpy as np
import pytest

from tests import common as T

pytestmark = pytest.mark.gpu
SEED = 0          # bench.py's seed: these are the benchmark's volumes


def _need_memory(gb):
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < gb * 2 ** 30:
        pytest.skip("needs %d GB of free HBM" % gb)


def _lattice(R, n=64):
    """n pixel indices spread over [0, R): one per R/n block, at a position inside the block that walks through every
    residue modulo 16 and 32 (a regular stride would only ever test lane 0 of every CTA tile)."""
    step = R // n
    k = np.arange(n)
    return k * step + (k * 5 + 3) % step


def _torch_done():
    """torch generated the volume on ITS stream; the samplers launch on the context's own non-blocking stream."""
    import torch
    torch.cuda.synchronize()


def _reference():
    from oracle import oracle as O
    if O.ref_available():
        S, G = O.load_ref()
        return S, G, "reference"
    return O, O, "oracle"


def lattice_rays(cam, args, ii, jj):
    """vp_pos / vp_dir (len(ii), len(jj), 3) of the plane-parallel rays of pixels ii x jj, computed as
    generate_vector_info_plane_parallel does (lenses.pyx:179-192)."""
    nx, ny = cam.resolution
    vp = np.asarray(args[0], "float64").ravel()
    vd = np.asarray(args[1], "float64").ravel()[:3]
    w = np.asarray(args[7], "float64")
    px = w[0] * np.asarray(ii, "float64") / (nx - 1.0) - w[0] / 2.0
    py = w[1] * np.asarray(jj, "float64") / (ny - 1.0) - w[1] / 2.0
    pos = (vp[0:3][None, None, :] * px[:, None, None] + vp[3:6][None, None, :] * py[None, :, None]) + vp[9:12][None, None, :]
    dirs = np.broadcast_to(vd, pos.shape)
    return np.ascontiguousarray(pos), np.ascontiguousarray(dirs), vd


def x_slabs(n, n_slabs, vd):
    """Cell ranges [x0, x1) of `n_slabs` x-slabs in the order a ray of direction vd meets them."""
    edges = np.linspace(0, n, n_slabs + 1).astype(int)
    out = [(int(edges[k]), int(edges[k + 1])) for k in range(n_slabs)]
    return out if vd[0] >= 0 else out[::-1]


def reference_on_lattice(kind, cam, args, ii, jj, n, slab_of, centred, tf=None, n_samples=10, n_slabs=8, threads=0):
    """The reference sampler of `kind` over the rays of pixels ii x jj through the n^3 volume, slab by slab.
    `slab_of(x0, x1)` returns the host fp64 array of cells [x0, x1) (vertices [x0, x1] when `centred`)."""
    S, G, label = _reference()
    pos, dirs, vd = lattice_rays(cam, args, ii, jj)
    img = np.zeros((len(ii), len(jj), 4))
    a = list(args)
    a[0], a[1], a[4] = pos, dirs, img
    kw = {"lens_type": "lattice", "zbuffer": np.ones(img.shape[:2])}
    if kind == "vr":
        s = S.VolumeRenderSampler(*a, tf, n_samples, **kw)
    elif kind == "iproj":
        s = S.InterpolatedProjectionSampler(*a, n_samples, **kw)
    else:
        s = S.ProjectionSampler(*a, n_samples, **kw)
    for x0, x1 in x_slabs(n, n_slabs, vd):
        data = slab_of(x0, x1)
        assert data.shape[0] == x1 - x0 + centred and data.dtype == np.float64
        pg = G.PartitionedGrid(0, [data], np.ones((x1 - x0, n, n), "uint8"), np.array([x0 / n, 0.0, 0.0]), np.array([x1 / n, 1.0, 1.0]),
                               np.array([x1 - x0, n, n], "int64"))
        s(pg, num_threads=threads)
        del pg, data
    return np.array(s.aimage, copy=True), label


def test_north_star_frame_matches_the_reference_on_a_ray_lattice(gpu):
    """The 1024^3 @ 2048^2 frame bench.py times (same volume, camera, tables, n_samples), against the reference."""
    import torch
    from bench import N_SAMPLES, bench_camera, bench_tf
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import synthetic
    N, R = 1024, 2048
    _need_memory(12)
    dev = torch.device("cuda", gpu.device)
    cam, tf = bench_camera(R), bench_tf()
    args, kw = T.sampler_args(cam)
    kw.pop("camera_data")
    s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, **kw)
    vol = synthetic.vertex_slab(N, [0, 0, 0], [N, N, N], SEED, dev)
    bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
    bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [N] * 3, keep=vol)
    bs.commit()
    _torch_done()
    s.cast_bricks(bs, np.arange(1, dtype=np.int64))
    got = np.array(s.aimage)
    ii, jj = _lattice(R), _lattice(R)[::-1].copy()
    want, label = reference_on_lattice("vr", cam, args, ii, jj, N, lambda x0, x1: vol[x0:x1 + 1].cpu().numpy().astype("float64"), 1,
                                       tf=tf, n_samples=N_SAMPLES)
    sub = got[np.ix_(ii, jj)]
    assert want[:, :, :3].max() > 0 and (want[:, :, :3].sum(axis=2) > 0).mean() > 0.3, label
    err = T.normalised_diff(sub, want, False)
    print("north star vs %s on a %dx%d ray lattice: normalised max diff %.3g" % (label, len(ii), len(jj), err))
    assert err <= 1e-4
    assert np.array_equal(sub[:, :, 3], want[:, :, 3])     # non-grey: alpha untouched (H9)


def test_cfg3_random_field_projection_matches_the_reference_on_a_ray_lattice(gpu):
    """BASELINE configs[2]: off-axis ProjectionSampler of the benchmark's random 1024^3 field at 2048^2."""
    import torch
    from bench import bench_camera
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import synthetic
    N, R = 1024, 2048
    _need_memory(8)
    dev = torch.device("cuda", gpu.device)
    cam = bench_camera(R)
    args, kw = T.sampler_args(cam)
    kw.pop("camera_data")
    s = S.ProjectionSampler(*args, 10, **kw)
    vol, pitch = synthetic.cell_slab(N, [0, 0, 0], [N, N, N], SEED, dev)
    bs = D.BrickSet(s._ctx, 1, 1, L.CELL_CENTRED, L.STORE_F32)
    bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [N] * 3, keep=vol, row_pitch=pitch)
    bs.commit()
    _torch_done()
    s.cast_bricks(bs, np.arange(1, dtype=np.int64))
    got = np.array(s.aimage)
    ii, jj = _lattice(R), _lattice(R)[::-1].copy()
    want, label = reference_on_lattice("proj", cam, args, ii, jj, N, lambda x0, x1: vol[x0:x1].cpu().numpy().astype("float64"), 0)
    sub = got[np.ix_(ii, jj)]
    assert want[:, :, 0].max() > 0.5 and (want[:, :, 0] > 0).mean() > 0.3
    err = np.abs(sub[:, :, 0] - want[:, :, 0]).max() / want[:, :, 0].max()
    print("cfg3 vs %s on a %dx%d ray lattice: max diff / max %.3g" % (label, len(ii), len(jj), err))
    assert err <= 1e-6
    # per pixel too, wherever the integral is not a sliver of the volume's edge
    big = want[:, :, 0] > 1e-3 * want[:, :, 0].max()
    assert np.abs(sub[:, :, 0][big] / want[:, :, 0][big] - 1.0).max() <= 1e-6
    assert np.all(got[:, :, 1:] == 0.0)


def test_cfg2_full_frame_matches_the_reference(gpu):
    """BASELINE configs[1] in full: 512^3 VolumeSource, ColorTransferFunction, off-axis 1024^2 -- every pixel."""
    import torch
    from bench import N_SAMPLES, bench_camera, bench_tf
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import synthetic
    N, R = 512, 1024
    _need_memory(4)
    dev = torch.device("cuda", gpu.device)
    cam, tf = bench_camera(R), bench_tf()
    args, kw = T.sampler_args(cam)
    kw.pop("camera_data")
    s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, **kw)
    vol = synthetic.vertex_slab(N, [0, 0, 0], [N, N, N], SEED, dev)
    bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
    bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [N] * 3, keep=vol)
    bs.commit()
    _torch_done()
    s.cast_bricks(bs, np.arange(1, dtype=np.int64))
    got = np.array(s.aimage)
    Sr, Gr, label = _reference()
    img = np.zeros((R, R, 4))
    a2, kw2 = T.sampler_args(cam, image=img)
    kw2.pop("camera_data")
    ref = Sr.VolumeRenderSampler(*a2, tf, N_SAMPLES, **kw2)
    data = vol.cpu().numpy().astype("float64")
    pg = Gr.PartitionedGrid(0, [data], np.ones((N, N, N), "uint8"), np.zeros(3), np.ones(3), np.array([N] * 3, "int64"))
    ref(pg, num_threads=0)
    want = np.array(ref.aimage)
    assert want[:, :, :3].max() > 0
    err = T.normalised_diff(got, want, False)
    print("cfg2 full frame vs %s: normalised max diff %.3g" % (label, err))
    assert err <= 1e-4


def test_cfg4_amr_frame_matches_the_reference_on_a_ray_lattice(gpu):
    """BASELINE configs[3] as bench.py times it (nested-box AMR, 16 384 grids / 14 848 kd leaves of 8^3 cells,
    PerspectiveLens at 2048^2, the whole view-ordered leaf list in one launch), against the reference casting a 64x64
    lattice of the same perspective rays through the same leaves in the same kd order, leaf by leaf.  The leaves reach
    the reference as the host mirror builds them (get_brick_data: the reference's own per-grid vertex centring),
    rounded to the fp32 the device stores."""
    from bench import AMR_WORKLOADS, N_SAMPLES, amr_camera, bench_tf
    from yt_b200 import _lib as L
    from yt_b200 import image_samplers as S
    from yt_b200 import loaders, synthetic
    from yt_b200.amr_kdtree import AMRKDTree
    R = 2048
    base, tile, levels = AMR_WORKLOADS["cfg4_amr_vr_15kbricks_persp_2048px"]
    grids, dd = synthetic.nested_amr_grids(base_cells=base, tile=tile, levels=levels, seed=SEED)
    ds = loaders.load_amr_grids(grids, dd)
    kd = AMRKDTree(ds, comm_rank=0, comm_size=1)
    kd.set_fields([ds.default_field], [False], True)
    cam, tf = amr_camera(R), bench_tf()
    args, kw = T.sampler_args(cam)
    s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, **kw)
    bricks = kd.device_bricks(s._ctx, L.STORE_F32)
    order = np.ascontiguousarray(kd.view_order(cam.lens.viewpoint), dtype=np.int64)
    assert len(order) > 14000
    s.cast_bricks(bricks, order)
    got = np.array(s.aimage)

    Sref, Gref, label = _reference()
    ii, jj = _lattice(R), _lattice(R)[::-1].copy()
    pos = np.ascontiguousarray(np.asarray(args[0], "float64")[np.ix_(ii, jj)])
    dirs = np.ascontiguousarray(np.asarray(args[1], "float64")[np.ix_(ii, jj)])
    img = np.zeros((len(ii), len(jj), 4))
    a = list(args)
    a[0], a[1], a[4] = pos, dirs, img
    ref = Sref.VolumeRenderSampler(*a, tf, N_SAMPLES, lens_type="lattice", zbuffer=np.ones(img.shape[:2]))
    n_leaves = 0
    for node in kd.tree.traverse(cam.lens.viewpoint):
        b = kd.get_brick_data(int(node))
        data = [np.ascontiguousarray(d.astype("float32").astype("float64")) for d in b.my_data]
        ref(Gref.PartitionedGrid(0, data, np.asarray(b.source_mask, "uint8"), np.asarray(b.LeftEdge, "float64"), np.asarray(b.RightEdge, "float64"),
                                 np.asarray(b.dims, "int64")), num_threads=0)
        n_leaves += 1
    assert n_leaves == len(order)
    want = np.array(ref.aimage)
    sub = got[np.ix_(ii, jj)]
    assert (want[:, :, :3].sum(axis=2) > 0).mean() > 0.3, label
    err = T.normalised_diff(sub, want, False)
    print("cfg4 vs %s on a %dx%d ray lattice over %d leaves: normalised max diff %.3g" % (label, len(ii), len(jj), n_leaves, err))
    assert err <= 1e-4
    assert np.array_equal(sub[:, :, 3], want[:, :, 3])


@pytest.mark.parametrize("opacity", [1.0, 30.0])
def test_grey_opacity_frames_match_the_reference_on_a_ray_lattice(opacity, gpu):
    """The north-star volume through the benchmark's table with grey opacity (and 30x more opaque: the case early ray
    termination is for), plain cast and terminated cast (eps 1e-6), against the reference: ~18 000 fp32 compositing
    steps per ray in either direction."""
    import torch
    from bench import N_SAMPLES, bench_camera, bench_tf
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import synthetic
    from yt_b200.volume_rendering.transfer_functions import ColorTransferFunction
    N, R = 1024, 2048
    _need_memory(12)
    dev = torch.device("cuda", gpu.device)
    cam, base = bench_camera(R), bench_tf()
    tf = ColorTransferFunction(base.x_bounds, base.nbins, grey_opacity=True)
    for c in range(4):
        tf.funcs[c].y[:] = base.funcs[c].y * (opacity if c == 3 else 1.0)
    args, kw = T.sampler_args(cam)
    kw.pop("camera_data")
    vol = synthetic.vertex_slab(N, [0, 0, 0], [N, N, N], SEED, dev)
    _torch_done()
    ii, jj = _lattice(R), _lattice(R)[::-1].copy()
    want, label = reference_on_lattice("vr", cam, args, ii, jj, N, lambda x0, x1: vol[x0:x1 + 1].cpu().numpy().astype("float64"), 1,
                                       tf=tf, n_samples=N_SAMPLES)
    for ert in (0.0, 1e-6):
        args, kw = T.sampler_args(cam)      # a fresh (zeroed) image: the sampler composites onto what it is given
        kw.pop("camera_data")
        s = S.VolumeRenderSampler(*args, tf, N_SAMPLES, early_termination=ert, **kw)
        bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
        bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [N] * 3, keep=vol)
        bs.commit()
        s.cast_bricks(bs, np.arange(1, dtype=np.int64))
        sub = np.array(s.aimage)[np.ix_(ii, jj)]
        err = T.normalised_diff(sub, want, True)
        print("grey opacity x%g, early termination %g vs %s: normalised max diff %.3g" % (opacity, ert, label, err))
        assert err <= 1e-4


def test_interpolated_projection_matches_the_reference_on_a_ray_lattice(gpu):
    """InterpolatedProjectionSampler over the north-star volume (bench workload iproj_1024c_2048px): fp32 cell sums on the
    corner-reuse walk, fp64 accumulation, against the reference's fp64 samples; bar rtol 1e-6 of the largest integral."""
    import torch
    from bench import N_SAMPLES, bench_camera
    from yt_b200 import _lib as L
    from yt_b200 import device as D
    from yt_b200 import image_samplers as S
    from yt_b200 import synthetic
    N, R = 1024, 2048
    _need_memory(12)
    dev = torch.device("cuda", gpu.device)
    cam = bench_camera(R)
    args, kw = T.sampler_args(cam)
    kw.pop("camera_data")
    s = S.InterpolatedProjectionSampler(*args, N_SAMPLES, **kw)
    vol = synthetic.vertex_slab(N, [0, 0, 0], [N, N, N], SEED, dev)
    bs = D.BrickSet(s._ctx, 1, 1, L.VERTEX_CENTRED, L.STORE_F32)
    bs.attach(0, [vol.data_ptr()], 0, [0.0] * 3, [1.0] * 3, [N] * 3, keep=vol)
    bs.commit()
    _torch_done()
    s.cast_bricks(bs, np.arange(1, dtype=np.int64))
    got = np.array(s.aimage)
    ii, jj = _lattice(R), _lattice(R)[::-1].copy()
    want, label = reference_on_lattice("iproj", cam, args, ii, jj, N, lambda x0, x1: vol[x0:x1 + 1].cpu().numpy().astype("float64"), 1,
                                       n_samples=N_SAMPLES)
    sub = got[np.ix_(ii, jj)]
    err = np.abs(sub[:, :, 0] - want[:, :, 0]).max() / np.abs(want[:, :, 0]).max()
    print("iproj vs %s on a %dx%d ray lattice: max diff / max %.3g" % (label, len(ii), len(jj), err))
    assert np.abs(want[:, :, 0]).max() > 0
    assert err <= 1e-6
