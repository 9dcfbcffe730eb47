"""PartitionedGrid -- host-side mirror of yt/utilities/lib/partitioned_grid.pyx:20-70.

Same constructor and public attributes (``my_data``, ``source_mask``, ``LeftEdge``,
``RightEdge``, ``parent_grid_id``; partitioned_grid.pxd:10-21).  The device copy is made
lazily by :meth:`upload` and cached, like ``node.data`` in amr_kdtree.py:326-327.
"""
import numpy as np

from . import _lib as L
from . import device as D


class PartitionedGrid:
    def __init__(self, parent_grid_id, data, mask, left_edge, right_edge, dims, n_fields=-1):
        self.parent_grid_id = int(parent_grid_id)
        self.LeftEdge = np.asarray(left_edge, dtype=np.float64)
        self.RightEdge = np.asarray(right_edge, dtype=np.float64)
        dims = np.asarray(dims)
        if dims.dtype != np.int64:
            # the reference's typed argument is np.int64_t[:] (partitioned_grid.pyx:33)
            raise ValueError("Buffer dtype mismatch, expected 'int64_t' but got '%s'" % dims.dtype)
        self.dims = dims
        if n_fields == -1:
            n_fields = len(data)
        self.n_fields = int(n_fields)
        self.my_data = data
        self.source_mask = mask
        for d in data:
            if not (isinstance(d, np.ndarray) and d.dtype == np.float64 and d.ndim == 3):
                raise ValueError("Buffer dtype mismatch, expected 3-D 'float64_t' brick data")
        self._dev = {}  # (device, storage, centring) -> 1-brick BrickSet

    @property
    def centring(self):
        """VERTEX_CENTRED if data is (dims+1), CELL_CENTRED if data is dims (ProjectionSampler)."""
        shp = tuple(self.my_data[0].shape)
        if shp == tuple(int(d) + 1 for d in self.dims):
            return L.VERTEX_CENTRED
        if shp == tuple(int(d) for d in self.dims):
            return L.CELL_CENTRED
        raise ValueError(f"brick data shape {shp} matches neither dims nor dims+1 ({tuple(self.dims)})")

    def upload(self, ctx=None, storage=L.STORE_F32):
        """Upload once per (device, storage); returns the owning 1-brick set."""
        ctx = ctx or D.get_context()
        cen = self.centring
        key = (ctx.device, storage, cen)
        if key not in self._dev:
            s = D.BrickSet(ctx, 1, self.n_fields, cen, storage)
            s.upload(0, self.my_data[: self.n_fields], self.source_mask, self.LeftEdge, self.RightEdge, self.dims)
            s.commit()
            self._dev[key] = s
        return self._dev[key]


def build_brickset(bricks, ctx=None, storage=L.STORE_F32):
    """One device brick table over many PartitionedGrids (device arrays are shared, not copied)."""
    ctx = ctx or D.get_context()
    if len(bricks) == 0:
        raise ValueError("no bricks")
    nf, cen = bricks[0].n_fields, bricks[0].centring
    out = D.BrickSet(ctx, len(bricks), nf, cen, storage)
    for i, b in enumerate(bricks):
        if b.n_fields != nf or b.centring != cen:
            raise ValueError("all bricks of a set need the same field count and centring")
        one = b.upload(ctx, storage)
        ptrs, mask = one.device_ptrs(0)
        out.attach(i, ptrs, mask, b.LeftEdge, b.RightEdge, b.dims, keep=one, row_pitch=one.pitch(0)[1])
    return out.commit()
