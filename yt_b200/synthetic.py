"""Synthetic uniform-grid volumes generated directly in HBM (torch is used for device buffers and
its Philox generator only -- plumbing, not the product).

The global field is defined plane by plane (``cells[x] = U[0,1)`` from a generator seeded with
``seed * 4099 + x``), so any rank can materialise exactly its own slab of the same global volume
(BASELINE.json configs 2, 3, 5 and the north-star frame).  Vertex centring follows the
reference's no-ghost rule (grid_patch.py:284-296 -> interpolators.pyx:177-227): interior vertices
are the mean of their neighbours along each axis, boundary vertices are linearly extrapolated.
"""
import numpy as np


def _torch():
    import torch
    return torch


def cell_planes(n, x0, x1, seed, device, ny=None, nz=None):
    """fp32 (x1-x0, ny, nz) cell-centred slab of the global random field."""
    torch = _torch()
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    out = torch.empty((x1 - x0, ny, nz), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    for x in range(x0, x1):
        g.manual_seed(int(seed) * 4099 + x)
        out[x - x0] = torch.rand((ny, nz), generator=g, device=device, dtype=torch.float32)
    return out


def _centre_axis(c, axis, n_global, c0, v0, v1):
    """Vertices [v0, v1] (inclusive, global indices) along `axis` from cells [c0, c0 + c.shape[axis])."""
    torch = _torch()
    idx = torch.arange(v0, v1 + 1, device=c.device)
    lo = (idx - 1).clamp(0, n_global - 2)
    lo = torch.where(idx >= n_global, torch.full_like(lo, n_global - 2), lo)
    # weight of the upper neighbour: 0.5 inside, -0.5 / 1.5 at the two domain faces
    w = torch.full(idx.shape, 0.5, dtype=c.dtype, device=c.device)
    w = torch.where(idx == 0, torch.full_like(w, -0.5), w)
    w = torch.where(idx == n_global, torch.full_like(w, 1.5), w)
    a = c.index_select(axis, lo - c0)
    b = c.index_select(axis, lo + 1 - c0)
    shape = [1, 1, 1]
    shape[axis] = -1
    w = w.reshape(shape)
    return a * (1.0 - w) + b * w


def vertex_slab(n, lo, hi, seed, device):
    """Vertex-centred fp32 block of the global n^3 field covering cells [lo[a], hi[a]) on each axis:
    shape (hi-lo+1) per axis.  Shared vertex planes are bit-identical between neighbouring blocks."""
    lo, hi = [int(v) for v in lo], [int(v) for v in hi]
    cx0, cx1 = max(lo[0] - 1, 0), min(hi[0] + 1, n)
    if n >= 2:
        cx0, cx1 = min(cx0, n - 2), max(cx1, 2)
    c = cell_planes(n, cx0, cx1, seed, device)
    v = _centre_axis(c, 0, n, cx0, lo[0], hi[0])
    del c
    v = _centre_axis(v, 1, n, 0, lo[1], hi[1])
    v = _centre_axis(v, 2, n, 0, lo[2], hi[2])
    return v.contiguous()


def padded_row(row_elems, elem_bytes=4):
    """Row pitch the library uses for arrays it owns: rows that are a multiple of 256 B get one 32-byte sector of
    padding (power-of-two row strides thrash the L1 sets, see DevBrick::pitch_y)."""
    return row_elems + 32 // elem_bytes if (row_elems * elem_bytes) % 256 == 0 else row_elems


def cell_slab(n, lo, hi, seed, device, pad=True):
    """Cell-centred fp32 block [lo, hi) of the global field (ProjectionSampler input).  Returns (tensor, row_pitch):
    the tensor is a view of shape (hi-lo) whose z-rows are `row_pitch` elements apart."""
    torch = _torch()
    c = cell_planes(n, int(lo[0]), int(hi[0]), seed, device)
    c = c[:, int(lo[1]):int(hi[1]), int(lo[2]):int(hi[2])]
    nz = c.shape[2]
    pitch = padded_row(nz) if pad else nz
    out = torch.zeros((c.shape[0], c.shape[1], pitch), dtype=c.dtype, device=c.device)
    out[:, :, :nz] = c
    return out[:, :, :nz], pitch


def host_cells(n, seed):
    """Small host-side fp64 volume for the CPU baseline sample (numpy generator; not the device field)."""
    return np.random.default_rng(seed).random((n, n, n))


def nested_amr_grids(base_cells=32, tile=8, levels=4, seed=0, refine_by=2, field="density"):
    """Host-side nested-box AMR hierarchy in ``load_amr_grids`` form (BASELINE.json config 4).

    Level 0 tiles the unit domain (``base_cells``^3 cells) with ``tile``^3-cell grids; every finer
    level refines the central half (per axis) of the previous level's region, again tiled by
    ``tile``^3-cell grids aligned to parent cell edges.  Data are U[0,1) * 2^level from
    ``default_rng(seed)`` drawn grid by grid."""
    rng = np.random.default_rng(seed)
    out = []
    lo, hi = 0.0, 1.0
    for lvl in range(levels):
        n = base_cells * refine_by ** lvl          # cells across the whole domain at this level
        dx = 1.0 / n
        c0, c1 = int(round(lo * n)), int(round(hi * n))
        for i in range(c0, c1, tile):
            for j in range(c0, c1, tile):
                for k in range(c0, c1, tile):
                    dims = [min(tile, c1 - i), min(tile, c1 - j), min(tile, c1 - k)]
                    out.append({"left_edge": np.array([i, j, k]) * dx,
                                "right_edge": (np.array([i, j, k]) + dims) * dx,
                                "level": lvl, "dimensions": np.array(dims, dtype="int64"),
                                field: rng.random(tuple(dims)) * 2.0 ** lvl + 0.05})
        q = (hi - lo) / 4.0
        lo, hi = lo + q, hi - q
    return out, [base_cells] * 3


class BoxDomain:
    """The few dataset attributes a Camera reads (camera.py:127-137), for a bare box with no fields."""

    def __init__(self, le=(0, 0, 0), re=(1, 1, 1)):
        self.domain_left_edge = np.array(le, dtype="float64")
        self.domain_right_edge = np.array(re, dtype="float64")
        self.domain_width = self.domain_right_edge - self.domain_left_edge
        self.domain_center = 0.5 * (self.domain_left_edge + self.domain_right_edge)
        self.ds = self

    def all_data(self):
        return self


def make_camera(lens_type="plane-parallel", resolution=(64, 64), position=None, focus=None, width=None, north=None, ds=None):
    """A Camera over `ds` (default: the unit box) with the reference's defaults -- the corner view of
    ``yt.create_scene`` -- unless position / focus / width / north are given."""
    from .volume_rendering.camera import Camera
    cam = Camera(None, ds or BoxDomain(), lens_type=lens_type)
    cam.resolution = resolution
    if north is not None:
        cam.north_vector = np.array(north, dtype="float64")
    if position is not None:
        cam.position = position
    if focus is not None:
        cam.focus = focus
    if width is not None:
        cam.width = width
    return cam
