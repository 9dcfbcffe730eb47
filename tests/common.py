"""Shared scenario builders for the parity tests: cameras, bricks and transfer functions.

Sampler arguments come from the product's own host mirror (yt_b200.volume_rendering), which
restates the reference's lens / camera code; the same argument tuple is then fed to the compiled
reference (oracle/_ref), the C oracle (oracle/ytray_oracle.c) and the CUDA samplers.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from yt_b200.volume_rendering.camera import Camera  # noqa: E402
from yt_b200.volume_rendering.transfer_functions import ColorTransferFunction, MultiVariateTransferFunction, TransferFunction  # noqa: E402


from yt_b200.synthetic import BoxDomain as FakeDS, make_camera  # noqa: E402,F401
from yt_b200.volume_rendering.utils import sampler_arguments as sampler_args  # noqa: E402,F401


def vertex_centre(cells):
    """No-ghost vertex centring of one grid: trilinear extrapolation from cell centres
    (grid_patch.py:284-296 -> interpolators.pyx:177-227 with rf=1, input_left=0.5)."""
    out = np.asarray(cells, dtype="float64")
    for ax in range(3):
        n = out.shape[ax]
        o = np.arange(n + 1, dtype="float64")
        r = o - 0.5
        i = np.clip(r.astype("int64"), 0, max(n - 2, 0))  # C cast truncates toward zero like <int>
        i[r < 0] = 0
        xp = r - i
        lo = np.take(out, i, axis=ax)
        hi = np.take(out, np.minimum(i + 1, n - 1), axis=ax)
        shape = [1, 1, 1]
        shape[ax] = n + 1
        xp = xp.reshape(shape)
        out = lo * (1.0 - xp) + hi * xp
    return np.ascontiguousarray(out)


def split_brick(data_v, mask, le, re, nsplit):
    """Split one vertex-centred brick into nsplit[0]*nsplit[1]*nsplit[2] cell-aligned sub-bricks
    (shared vertex planes duplicated, like amr_kdtree.py:362-365)."""
    dims = np.array(mask.shape)
    le, re = np.asarray(le, "float64"), np.asarray(re, "float64")
    dds = (re - le) / dims
    edges = [np.linspace(0, dims[a], nsplit[a] + 1).astype(int) for a in range(3)]
    out = []
    for i in range(nsplit[0]):
        for j in range(nsplit[1]):
            for k in range(nsplit[2]):
                lo = np.array([edges[0][i], edges[1][j], edges[2][k]])
                hi = np.array([edges[0][i + 1], edges[1][j + 1], edges[2][k + 1]])
                d = [np.ascontiguousarray(f[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1]) for f in data_v]
                m = np.ascontiguousarray(mask[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]])
                out.append((d, m, le + lo * dds, le + hi * dds, (hi - lo).astype("int64")))
    return out


def view_order(bricks_geom, viewpoint):
    """Back-to-front order of axis-aligned boxes on a lattice w.r.t. a viewpoint: farthest first.
    For a regular split this equals the kd traversal (amr_kdtools.pyx:708-755) up to ties that do
    not overlap in any ray."""
    centres = np.array([0.5 * (g[0] + g[1]) for g in bricks_geom])
    d = np.linalg.norm(centres - np.asarray(viewpoint)[None, :], axis=1)
    return np.argsort(-d, kind="stable")


def colour_tf(bounds=(0.0, 1.0), nbins=256, grey=False, n_gauss=4, seed=1, height=1.0):
    rng = np.random.RandomState(seed)
    tf = ColorTransferFunction(bounds, nbins, grey_opacity=grey)
    span = bounds[1] - bounds[0]
    for k in range(n_gauss):
        loc = bounds[0] + span * (k + 0.5) / n_gauss
        col = list(rng.uniform(0.1, 1.0, 3)) + [height * 10.0 ** rng.uniform(-1.5, 0.0)]
        tf.add_gaussian(loc, (0.03 * span) ** 2, col)
    return tf


def multivariate_tf(bounds0=(0.0, 1.0), bounds1=(0.0, 2.0), grey=False):
    """Two-field, five-table transfer function exercising weight_field / weight_table chains."""
    mv = MultiVariateTransferFunction(grey_opacity=grey)
    t = []
    for k, (b, nb) in enumerate([(bounds0, 64), (bounds1, 33), (bounds0, 128), (bounds1, 17), (bounds0, 50)]):
        f = TransferFunction(b, nb)
        f.add_gaussian(b[0] + (0.2 + 0.15 * k) * (b[1] - b[0]), (0.1 * (b[1] - b[0])) ** 2, 0.5 + 0.1 * k)
        f.add_line((b[0], 0.05), (b[1], 0.3))
        t.append(f)
    mv.add_field_table(t[0], 0, weight_table_id=3)
    mv.add_field_table(t[1], 1, weight_field_id=0)
    mv.add_field_table(t[2], 0, weight_field_id=1, weight_table_id=1)
    mv.add_field_table(t[3], 1)
    mv.add_field_table(t[4], 0, weight_table_id=0)
    mv.link_channels(0, 0)
    mv.link_channels(2, 1)
    mv.link_channels(4, 2)
    mv.link_channels(3, 3)
    return mv


from yt_b200.metrics import normalised_diff  # noqa: E402,F401


# ---------------------------------------------------------------------------------------------------
# scenarios: small cases every implementation (reference, C oracle, CUDA) is run on
# ---------------------------------------------------------------------------------------------------
class Scenario:
    def __init__(self, name, kind, cam, bricks, order, tf=None, n_samples=10, zbuffer=None, image0=None, grey=False):
        self.name, self.kind, self.cam = name, kind, cam
        self.bricks = bricks          # list of (data list, mask, le, re, dims[int64])
        self.order = list(order)
        self.tf, self.n_samples, self.zbuffer, self.image0, self.grey = tf, n_samples, zbuffer, image0, grey


def _rand_cells(n, seed, lo=0.0, hi=1.0):
    if np.ndim(n) == 0:
        n = (n, n, n)
    return np.random.RandomState(seed).uniform(lo, hi, size=tuple(n))


def zero_padded(d):
    """The same array inside a larger zeroed buffer.  A ray that enters a brick exactly through an edge makes the
    reference sample an out-of-range cell over a zero-length segment (grid_traversal.pyx:104-138): it reads past the
    array.  The plain samplers are immune (dt = 0 times a finite table value), the lit one multiplies by a gradient of
    whatever it read -- NaN or inf garbage turns the pixel NaN.  With zeros around the data the CPU implementations are
    deterministic (and equal to any finite garbage); the CUDA cast never touches that memory."""
    d = np.ascontiguousarray(d, dtype="float64")
    buf = np.zeros(3 * d.size + 192)
    view = buf[d.size + 64:2 * d.size + 64].reshape(d.shape)
    view[...] = d
    return view


def run_scenario(sc, sampler_mod, grid_mod, num_threads=1, sampler_kwargs=None):
    """Drive one implementation exactly as render_source.py:594-599 drives the reference."""
    img = np.zeros((sc.cam.resolution[0], sc.cam.resolution[1], 4), "float64") if sc.image0 is None else sc.image0.copy()
    args, kwargs = sampler_args(sc.cam, zbuffer=None if sc.zbuffer is None else sc.zbuffer.copy(), image=img)
    if sc.cam.lens.lens_type != "perspective" or sc.kind != "vr":
        # only new_volume_render_sampler forwards camera_data (utils.py:76-77)
        kwargs.pop("camera_data")
    if sampler_kwargs:
        kwargs.update(sampler_kwargs)
    if sc.kind == "vr":
        s = sampler_mod.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs)
    elif sc.kind == "iproj":
        s = sampler_mod.InterpolatedProjectionSampler(*args, sc.n_samples, **kwargs)
    else:
        s = sampler_mod.ProjectionSampler(*args, sc.n_samples, **kwargs)
    grids = [grid_mod.PartitionedGrid(i, [zero_padded(d) for d in b[0]], b[1], np.array(b[2], "float64"), np.array(b[3], "float64"),
                                      np.array(b[4], "int64")) for i, b in enumerate(sc.bricks)]
    for i in sc.order:
        s(grids[i], num_threads=num_threads)
    return np.array(s.aimage, copy=True)


def scenarios(small=True):
    """The parity matrix.  Sizes are chosen so the CPU oracle finishes each in well under a second."""
    out = []
    N, R = (16, 48) if small else (32, 96)
    unit = (np.zeros(3), np.ones(3))
    ones_mask = lambda n: np.ones((n, n, n) if np.ndim(n) == 0 else tuple(n), "uint8")  # noqa: E731

    # 1. ProjectionSampler, corner view, one cell-centred brick, two fields
    cam = make_camera(resolution=(R, R))
    out.append(Scenario("proj_corner", "proj", cam,
                        [([_rand_cells(N, 1), _rand_cells(N, 2, 0, 5)], ones_mask(N), *unit, [N] * 3)], [0]))
    # 2. ProjectionSampler, axis-aligned view: rays run exactly along cell faces (tie / edge cases H1-H4)
    cam = make_camera(resolution=(N + 1, N + 1), position=[0.5, 0.5, 2.0], width=[1.0, 1.0, 3.0])
    out.append(Scenario("proj_on_axis_edges", "proj", cam, [([_rand_cells(N, 3)], ones_mask(N), *unit, [N] * 3)], [0]))
    # 3. off-axis normal, non-cubic brick, off-centre domain, z-buffer cutting rays short
    cam = make_camera(resolution=(R, R + 7), position=[1.3, 0.2, 1.9], focus=[0.4, 0.55, 0.5], width=[1.2, 1.4, 2.5])
    dims = (N, N + 3, N - 5)
    zb = np.random.RandomState(4).uniform(-0.2, 1.3, size=(R, R + 7))
    out.append(Scenario("proj_offaxis_zbuf", "proj", cam,
                        [([_rand_cells(dims, 5)], ones_mask(dims), np.array([0.1, 0.0, 0.2]), np.array([0.9, 1.0, 0.7]), list(dims))],
                        [0], zbuffer=zb))
    # 4. InterpolatedProjectionSampler, 3 fields + a 4th that must be ignored, n_samples 5
    cam = make_camera(resolution=(R, R), position=[1.0, 0.9, 0.3])
    data = [vertex_centre(_rand_cells(N, 10 + k)) for k in range(4)]
    out.append(Scenario("iproj_multi", "iproj", cam, [(data, ones_mask(N), *unit, [N] * 3)], [0], n_samples=5))
    # 5. InterpolatedProjectionSampler one field
    out.append(Scenario("iproj_single", "iproj", make_camera(resolution=(R, R)),
                        [([vertex_centre(_rand_cells(N, 20))], ones_mask(N), *unit, [N] * 3)], [0]))
    # 6. VolumeRenderSampler, ColorTransferFunction (the default path), non-grey
    v = vertex_centre(_rand_cells(N, 30))
    out.append(Scenario("vr_colour", "vr", make_camera(resolution=(R, R)),
                        [([v], ones_mask(N), *unit, [N] * 3)], [0], tf=colour_tf(height=40.0)))
    # 7. same, grey opacity, with an initial image (opaque sources) and z-buffer
    rng = np.random.RandomState(31)
    img0 = rng.uniform(0, 0.3, size=(R, R, 4))
    zb = rng.uniform(0.3, 1.2, size=(R, R))
    out.append(Scenario("vr_colour_grey_zbuf", "vr", make_camera(resolution=(R, R), position=[0.2, 1.4, 1.1]),
                        [([v], ones_mask(N), *unit, [N] * 3)], [0], tf=colour_tf(grey=True, height=400.0), zbuffer=zb,
                        image0=img0, grey=True))
    # 8. masked cells + values outside the table bounds + NaN and 0 vertices (H7, H10)
    v2 = vertex_centre(_rand_cells(N, 32, -0.2, 1.2))
    v2[3:5, 3:5, :] = np.nan
    v2[8:10, :, 2:4] = 0.0
    m = (np.random.RandomState(33).uniform(size=(N, N, N)) > 0.3).astype("uint8")
    m[0, 0, 0] = 2
    out.append(Scenario("vr_mask_nan", "vr", make_camera(resolution=(R, R), position=[1.5, 0.1, 0.7]),
                        [([v2], m, *unit, [N] * 3)], [0], tf=colour_tf(height=300.0)))
    # 9. multi-field multivariate transfer function (weight_field / weight_table chains, H8)
    va, vb = vertex_centre(_rand_cells(N, 40)), vertex_centre(_rand_cells(N, 41, 0.0, 2.0))
    out.append(Scenario("vr_multivariate", "vr", make_camera(resolution=(R, R)),
                        [([va, vb], ones_mask(N), *unit, [N] * 3)], [0], tf=multivariate_tf(), n_samples=4))
    out.append(Scenario("vr_multivariate_grey", "vr", make_camera(resolution=(R, R), position=[0.1, 0.1, 1.6]),
                        [([va, vb], ones_mask(N), *unit, [N] * 3)], [0], tf=multivariate_tf(grey=True), n_samples=3, grey=True))
    # 10. eight sub-bricks in view order, plane-parallel (multi-brick list + footprint culling)
    sub = split_brick([v], np.ones((N, N, N), "uint8"), *unit, (2, 2, 2))
    cam = make_camera(resolution=(R, R), position=[1.2, 0.8, 1.1])
    order = view_order([(b[2], b[3]) for b in sub], cam.lens.viewpoint)
    out.append(Scenario("vr_8bricks_pp", "vr", cam, sub, order, tf=colour_tf(height=100.0)))
    # 11. perspective lens, camera outside, grey (order matters: H14), 8 sub-bricks
    cam = make_camera("perspective", resolution=(R, R), position=[1.6, 1.4, 1.2], width=[0.6, 0.6, 1.0])
    order = view_order([(b[2], b[3]) for b in sub], cam.lens.viewpoint)
    out.append(Scenario("vr_8bricks_perspective_grey", "vr", cam, sub, order, tf=colour_tf(grey=True, height=60.0), grey=True))
    # 12. perspective lens with the camera INSIDE the volume (H6)
    cam = make_camera("perspective", resolution=(R, R), position=[0.55, 0.45, 0.6], focus=[0.1, 0.2, 0.0], width=[0.8, 0.8, 0.5])
    order = view_order([(b[2], b[3]) for b in sub], cam.lens.viewpoint)
    out.append(Scenario("vr_perspective_inside", "vr", cam, sub, order, tf=colour_tf(height=30.0)))
    # 13. perspective projection samplers use the full image (calculate_extent_null)
    cam = make_camera("perspective", resolution=(R, R - 6), position=[1.9, 1.2, 0.4], width=[0.7, 0.7, 1.2])
    out.append(Scenario("proj_perspective", "proj", cam, [([_rand_cells(N, 50)], ones_mask(N), *unit, [N] * 3)], [0]))
    # 14. a brick the rays miss entirely + one that only some rays graze
    cam = make_camera(resolution=(R, R), width=[0.6, 0.6, 1.5])
    out.append(Scenario("vr_miss_and_graze", "vr", cam,
                        [([v], ones_mask(N), np.array([5.0, 5.0, 5.0]), np.array([6.0, 6.0, 6.0]), [N] * 3),
                         ([v], ones_mask(N), np.array([0.35, 0.1, 0.2]), np.array([0.6, 0.5, 0.9]), [N] * 3)],
                        [0, 1], tf=colour_tf(height=200.0)))
    # 15. 3x3 image and a single-cell brick
    cam = make_camera(resolution=(3, 3))
    out.append(Scenario("tiny", "vr", cam, [([np.arange(8, dtype="float64").reshape(2, 2, 2) / 8.0], np.ones((1, 1, 1), "uint8"),
                                            *unit, [1, 1, 1])], [0], tf=colour_tf(height=50.0)))
    return out


# ---------------------------------------------------------------------------------------------------
# octree scenarios (cast_through_octree, image_samplers.pyx:202-299)
# ---------------------------------------------------------------------------------------------------
class OctreeScenario:
    def __init__(self, name, kind, cam, depth, ipos, lvl, cells, tf=None, n_samples=10, image0=None, grey=False):
        self.name, self.kind, self.cam, self.depth = name, kind, cam, depth
        self.ipos, self.lvl, self.cells = ipos, lvl, cells      # int32 (n,3), int32 (n), float64 (n,14)
        self.tf, self.n_samples, self.image0, self.grey = tf, n_samples, image0, grey


def random_octree(depth, max_level, p_refine, seed, holes=0.0):
    """Leaves of a randomly refined octree over the unit cube: (ipos int32 (n,3) as octree_raytracing.py:35 builds
    them -- floor(centre * 2**depth) --, level int32 (n), centre (n,3), dx (n)).  `holes`: fraction of leaves dropped,
    so some children are never added to the tree (absent nodes, _octree_raytracing.hpp:273-274)."""
    rng = np.random.RandomState(seed)
    leaves = []

    def refine(i, j, k, level):
        if level < max_level and (level < 1 or rng.uniform() < p_refine):
            for di in range(2):
                for dj in range(2):
                    for dk in range(2):
                        refine(2 * i + di, 2 * j + dj, 2 * k + dk, level + 1)
        elif rng.uniform() >= holes:
            leaves.append((i, j, k, level))

    refine(0, 0, 0, 0)
    ijk = np.array([l[:3] for l in leaves], dtype="float64")
    lvl = np.array([l[3] for l in leaves], dtype="int32")
    dx = 1.0 / (1 << lvl).astype("float64")
    centre = (ijk + 0.5) * dx[:, None]
    ipos = np.floor(centre * (1 << depth)).astype("int64").astype("int32")
    return ipos, lvl, centre, dx


def octree_cells(centre, dx, seed, lo=0.05, hi=0.95):
    """The (n, 14) payload of OctreeVolumeSource.render (render_source.py:669-680): 8 vertex values of a smooth field
    plus noise (x-major within the cell, as set_fields leaves them), then left and right edge."""
    rng = np.random.RandomState(seed)
    n = len(dx)
    out = np.empty((n, 14), "float64")
    c = 0
    for a in range(2):
        for b in range(2):
            for k in range(2):
                v = centre + (np.array([a, b, k]) - 0.5)[None, :] * dx[:, None]
                f = 0.5 + 0.3 * np.sin(5.0 * v[:, 0]) * np.cos(4.0 * v[:, 1]) + 0.15 * np.sin(7.0 * v[:, 2])
                out[:, c] = np.clip(f + rng.uniform(-0.03, 0.03, n), lo, hi)
                c += 1
    out[:, 8:11] = centre - dx[:, None] / 2
    out[:, 11:14] = centre + dx[:, None] / 2
    return out


def octree_scenarios():
    out = []
    R = (40, 36)
    ipos, lvl, centre, dx = random_octree(depth=5, max_level=4, p_refine=0.45, seed=3)
    cells = octree_cells(centre, dx, 5)
    cam_a = make_camera("perspective", R, position=[1.7, 1.5, 1.3], width=[0.8, 0.8, 1.0])     # all directions negative
    cam_b = make_camera("perspective", R, position=[-0.6, 1.4, -0.9], width=[0.9, 0.9, 1.0])   # mixed signs
    cam_c = make_camera("perspective", R, position=[0.52, 0.47, 0.55], focus=[0.9, 0.1, 0.2], width=[1.2, 1.2, 1.0])  # inside
    rng = np.random.RandomState(9)
    out.append(OctreeScenario("oct_vr_colour", "vr", cam_a, 5, ipos, lvl, cells, tf=colour_tf(n_gauss=4)))
    out.append(OctreeScenario("oct_vr_grey_image0", "vr", cam_b, 5, ipos, lvl, cells, tf=colour_tf(grey=True, n_gauss=3, seed=2),
                              n_samples=7, image0=rng.uniform(0, 0.2, R + (4,)), grey=True))
    out.append(OctreeScenario("oct_vr_inside", "vr", cam_c, 5, ipos, lvl, cells, tf=colour_tf(n_gauss=5, seed=4), n_samples=4))
    out.append(OctreeScenario("oct_iproj", "iproj", cam_b, 5, ipos, lvl, cells, n_samples=5))
    out.append(OctreeScenario("oct_proj", "proj", cam_a, 5, ipos, lvl, cells))
    # a tree with absent children and a deeper maximum depth than any cell uses
    ipos2, lvl2, centre2, dx2 = random_octree(depth=7, max_level=5, p_refine=0.35, seed=8, holes=0.2)
    cells2 = octree_cells(centre2, dx2, 6)
    out.append(OctreeScenario("oct_vr_holes", "vr", cam_b, 7, ipos2, lvl2, cells2, tf=colour_tf(n_gauss=4, seed=7)))
    return out


def run_octree_scenario(sc, sampler_mod, grid_mod, octree_cls, num_threads=1, sampler_kwargs=None):
    """Drive one implementation as OctreeVolumeSource.render does (render_source.py:667-686)."""
    nx, ny = sc.cam.resolution
    img = np.zeros((nx, ny, 4), "float64") if sc.image0 is None else sc.image0.copy()
    args, kwargs = sampler_args(sc.cam, image=img)
    args[8] = "Octree"
    if sc.kind != "vr":
        kwargs.pop("camera_data")
    kwargs.update(sampler_kwargs or {})
    if sc.kind == "vr":
        s = sampler_mod.VolumeRenderSampler(*args, sc.tf, sc.n_samples, **kwargs)
    elif sc.kind == "iproj":
        s = sampler_mod.InterpolatedProjectionSampler(*args, sc.n_samples, **kwargs)
    else:
        s = sampler_mod.ProjectionSampler(*args, sc.n_samples, **kwargs)
    tree = octree_cls(np.zeros(3), np.ones(3), sc.depth)
    tree.add_nodes(sc.ipos, sc.lvl, np.arange(len(sc.lvl), dtype=np.int32))
    n = len(sc.lvl)
    dt = np.ascontiguousarray(sc.cells).reshape(1, n, 14, 1)
    mask = np.full(dt.shape[1:], 1, dtype=np.uint8)
    dims = np.array([1, 1, 1], dtype="int64")
    pg = grid_mod.PartitionedGrid(0, dt, mask, np.ascontiguousarray(sc.cells[:, 8:11]).ravel(),
                                  np.ascontiguousarray(sc.cells[:, 11:14]).ravel(), dims, n_fields=1)
    s(pg, oct=tree, num_threads=num_threads)
    return np.array(s.aimage, copy=True)
