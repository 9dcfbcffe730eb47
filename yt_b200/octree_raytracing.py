"""Octree ray tracing -- host mirror of yt/utilities/lib/_octree_raytracing.pyx:15-34 (``_OctreeRayTracing``) and of
the array part of yt/utilities/lib/octree_raytracing.py:9-41 (``OctreeRayTracing``).

The tree (``Octree<int>`` of _octree_raytracing.hpp) is built by libytb200.so on the host and flattened for the
device; rays are cast through it by ``ImageSampler.cast_through_octree`` (image_samplers.pyx:202-299), i.e. by
``sampler(pg, oct=tree)`` on a sampler whose ``volume_method`` is ``"Octree"``.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib as L
from . import device as D


class _OctreeRayTracing:
    """``_OctreeRayTracing(LE, RE, depth)`` + ``add_nodes(ipos, lvl, key)`` (same argument types: int32 views)."""

    def __init__(self, LE, RE, depth, context=None):
        if not isinstance(LE, np.ndarray) or not isinstance(RE, np.ndarray):
            raise TypeError("Argument 'LE' has incorrect type (expected numpy.ndarray)")
        le = np.ascontiguousarray(LE, dtype=np.float64)
        re = np.ascontiguousarray(RE, dtype=np.float64)
        self._ctx = context or D.get_context()
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.ytb200_octree_create(self._ctx.handle, int(depth), le.ctypes.data_as(L.dp), re.ctypes.data_as(L.dp),
                                         C.byref(h)))
        self.handle = h
        self.depth = int(depth)
        self._fin = weakref.finalize(self, lib.ytb200_octree_destroy, h)

    def add_nodes(self, ipos, lvl, key):
        for name, a, nd in (("ipos_view", ipos, 2), ("lvl_view", lvl, 1), ("key", key, 1)):
            a = np.asarray(a)
            if a.dtype != np.int32:   # int[:, :] / int[:] typed memoryviews
                raise ValueError("Buffer dtype mismatch, expected 'int' but got '%s'" % a.dtype)
            if a.ndim != nd:
                raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (nd, a.ndim))
        ipos = np.ascontiguousarray(ipos)
        lvl = np.ascontiguousarray(lvl)
        key = np.ascontiguousarray(key)
        L.check(L.load().ytb200_octree_add_nodes(self.handle, len(key), ipos.ctypes.data_as(L.i32p),
                                                 lvl.ctypes.data_as(L.i32p), key.ctypes.data_as(L.i32p)))

    @property
    def node_count(self):
        return int(L.load().ytb200_octree_node_count(self.handle))

    def set_cells(self, cells, token=None):
        """Upload the (n_cells, 14) payload of OctreeVolumeSource.render's PartitionedGrid.  A bare call uploads every
        time: the reference reads the host array at each cast, so in-place edits between casts must be seen.  A caller
        that OWNS the array and knows it unchanged (the scene's octree source, which assembles it from the volume's
        field data) passes a `token`: while the token is the one of the last upload the cells stay resident (75 MB per
        frame for a 675 k-cell tree: r01 e2e 29 ms against a 12.5 ms kernel)."""
        if token is not None and token == getattr(self, "_cells_token", None):
            return
        flat = np.ascontiguousarray(cells, dtype=np.float64).reshape(-1, 14)
        L.check(L.load().ytb200_octree_set_cells(self.handle, flat.shape[0], flat.ctypes.data_as(L.dp)))
        self._cells_token = token


class OctreeRayTracing:
    """octree_raytracing.py:9-41 for cells given as arrays (the reference reads them from a yt data object):
    ``xyz`` cell centres in unitary coordinates, ``grid_level`` their levels, ``min_level`` / ``max_level`` of the
    dataset.  ``self.octree`` is the ``_OctreeRayTracing``; ``self.data`` (8, n_cells) is set by the caller the way
    ``set_fields`` leaves it (vertex values, x-major within a cell)."""

    def __init__(self, xyz, grid_level, min_level, max_level, context=None):
        LE = np.array([0, 0, 0], dtype=np.float64)
        RE = np.array([1, 1, 1], dtype=np.float64)
        lvl_min = int(min_level) + 1
        depth = lvl_min + int(max_level) + 1
        self.octree = _OctreeRayTracing(LE, RE, depth, context=context)
        xyz = np.asarray(xyz, dtype=np.float64)
        lvl = np.asarray(grid_level).astype("int64", copy=False) + lvl_min
        ipos = np.floor(xyz * (1 << depth)).astype("int64")
        self.octree.add_nodes(ipos.astype(np.int32), lvl.astype(np.int32), np.arange(len(ipos), dtype=np.int32))
        self.data = None
