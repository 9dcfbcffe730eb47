"""Render sources -- yt/visualization/volume_rendering/render_source.py.

``VolumeSource`` / ``KDTreeVolumeSource`` / ``OctreeVolumeSource`` (209-693) with the reference's properties, setters and
``render``/``finalize_image`` contract.  The brick loop of ``render`` (594-597,
``for brick in volume.traverse(viewpoint): sampler(brick)``) is handed to the device as ONE launch:
all kd-tree leaves are kept resident as a brick table and only the view order is recomputed per
frame (``batched=True``, the default); ``batched=False`` drives the sampler brick by brick exactly
like the reference does.
"""
import abc
import os

import numpy as np

from .. import _lib as _L
from ..amr_kdtree import AMRKDTree
from ..loaders import OctreeIndex
from ..octree_raytracing import OctreeRayTracing
from ..partitioned_grid import PartitionedGrid
from .transfer_function_helper import TransferFunctionHelper
from .transfer_functions import ColorTransferFunction, ProjectionTransferFunction, TransferFunction
from .utils import data_source_or_all, new_interpolated_projection_sampler, new_projection_sampler, new_volume_render_sampler
from .zbuffer_array import ZBuffer


class RenderSource(abc.ABC):
    """Anything a Scene can render: ``render(camera, zbuffer)`` and ``_validate()``."""
    volume_method = None

    def __init__(self):
        self.opaque = False
        self.zbuffer = None

    @abc.abstractmethod
    def render(self, camera, zbuffer=None):
        ...

    @abc.abstractmethod
    def _validate(self):
        ...


class OpaqueSource(RenderSource):
    """Sources that paint the ZBuffer before the volume is cast."""

    def __init__(self):
        super().__init__()
        self.opaque = True

    def set_zbuffer(self, zbuffer):
        self.zbuffer = zbuffer


def create_volume_source(data_source, field):
    """render_source.py:170-190: a kd-tree source for grid datasets, an octree source for octree datasets."""
    data_source = data_source_or_all(data_source)
    if isinstance(data_source.ds.index, OctreeIndex):
        return OctreeVolumeSource(data_source, field)
    if not hasattr(data_source.ds.index, "grids"):
        raise NotImplementedError("Cannot render datasets that are neither grid-based nor octrees")
    return KDTreeVolumeSource(data_source, field)


_SAMPLER_FACTORY = {          # (sampler_type, interpolated) -> factory of utils.py:55-139
    ("volume-render", True): new_volume_render_sampler, ("volume-render", False): new_volume_render_sampler,
    ("projection", True): new_interpolated_projection_sampler, ("projection", False): new_projection_sampler,
}
_TF_TYPES = (TransferFunction, ColorTransferFunction, ProjectionTransferFunction, type(None))


class VolumeSource(RenderSource, abc.ABC):
    """State of one volume to be cast: data source, field, log / ghost-zone / weight choices, transfer function and
    the brick decomposition ("volume") they imply.  Same attributes, setters and invalidation rules as the reference
    (render_source.py:209-536): changing the field, the log flag, the ghost-zone flag or the weight field drops the
    decomposition (it is rebuilt, with the new fields, at the next render); a ProjectionTransferFunction switches the
    source to the projection samplers with a linear field and no ghost zones."""
    _image = None
    data_source = None

    def __init__(self, data_source, field):
        super().__init__()
        self.data_source = data_source_or_all(data_source)
        self._field = self.data_source._determine_fields(field)[0]
        self._log_field = self.data_source.ds.field_info[self._field].take_log
        self._use_ghost_zones = False
        self._weight_field = None
        self._transfer_function = None
        self._volume, self._volume_valid = None, False
        self.current_image = None
        self.check_nans = False
        self.num_threads, self.num_samples = 0, 10
        self.sampler_type = "volume-render"
        self.batched = True          # ours: all bricks in one launch (False: sampler(brick) per brick, like the reference loop)
        self.tfh = TransferFunctionHelper(self.data_source.pf)
        self.tfh.set_field(self._field)

    # -- invalidation -----------------------------------------------------------------------------------
    def _drop_volume(self):
        """What the reference's @invalidate_volume does after a setter (render_source.py:66-81)."""
        if isinstance(self._transfer_function, ProjectionTransferFunction):
            self.sampler_type = "projection"
            self._log_field = False
            self._use_ghost_zones = False
        self._volume = None
        self._volume_valid = False

    def _ensure_volume_fields(self):
        """What @validate_volume does before a render (render_source.py:84-103)."""
        if not self._volume_valid:
            fields, logs = [self.field], [self.log_field]
            if self.weight_field is not None:
                fields.append(self.weight_field)
                logs.append(self.log_field)
            self.volume.set_fields(fields, logs, no_ghost=(not self.use_ghost_zones))
        self._volume_valid = True

    # -- transfer function ------------------------------------------------------------------------------
    @property
    def transfer_function(self):
        if self._transfer_function is None:
            if self.tfh.tf is None:      # build the default: 10 layers over the field's extrema
                self.tfh.set_field(self.field)
                self.tfh.set_log(self.log_field)
                self.tfh.build_transfer_function()
                self.tfh.setup_default()
            self._transfer_function = self.tfh.tf
        return self._transfer_function

    @transfer_function.setter
    def transfer_function(self, value):
        self.tfh.tf = None
        if not isinstance(value, _TF_TYPES):
            raise RuntimeError("transfer_function not a valid type, received object of type %s" % type(value))
        if isinstance(value, ProjectionTransferFunction):
            self.sampler_type = "projection"
            if self._volume is not None:
                self._volume_valid = False
        self._transfer_function = value

    # -- the brick decomposition ------------------------------------------------------------------------
    @property
    def volume(self):
        return self._get_volume()

    @volume.setter
    def volume(self, value):
        assert isinstance(value, AMRKDTree)
        self._field, self._log_field = value.fields, value.log_fields
        self._volume = value
        assert self._volume_valid

    @volume.deleter
    def volume(self):
        self._volume = None

    @abc.abstractmethod
    def _get_volume(self):
        ...

    # -- what is rendered -------------------------------------------------------------------------------
    @property
    def field(self):
        return self._field

    @field.setter
    def field(self, value):
        found = self.data_source._determine_fields(value)
        if len(found) > 1:
            raise RuntimeError("VolumeSource.field can only be a single field but received multiple fields: %s" % found)
        new = found[0]
        if new != self._field:       # another field: its own log default, and new table bounds
            self._log_field = self.data_source.ds.field_info[new].take_log
            self.tfh.bounds = None
        self._field = new
        self.transfer_function = None
        self.tfh.set_field(new)
        self.tfh.set_log(self._log_field)
        self._drop_volume()

    @property
    def log_field(self):
        return self._log_field

    @log_field.setter
    def log_field(self, value):
        self.transfer_function = None
        self.tfh.set_log(value)
        self._log_field = value
        self._drop_volume()

    @property
    def use_ghost_zones(self):
        return self._use_ghost_zones

    @use_ghost_zones.setter
    def use_ghost_zones(self, value):
        self._use_ghost_zones = value
        self._drop_volume()

    @property
    def weight_field(self):
        return self._weight_field

    @weight_field.setter
    def weight_field(self, value):
        self._weight_field = value
        self._drop_volume()

    # chainable setters of the reference API (render_source.py:398-476)
    def _chain(attr):      # noqa: N805 -- helper evaluated at class-body time
        def setter(self, value):
            setattr(self, attr, value)
            return self
        return setter

    set_transfer_function = _chain("transfer_function")
    set_volume = _chain("volume")
    set_field = _chain("field")
    set_log = _chain("log_field")
    set_weight_field = _chain("weight_field")
    set_use_ghost_zones = _chain("use_ghost_zones")
    del _chain

    # -- sampler ----------------------------------------------------------------------------------------
    def _validate(self):
        if self.data_source is None:
            raise RuntimeError("Data source not initialized")

    def set_sampler(self, camera, interpolated=True):
        factory = _SAMPLER_FACTORY.get((self.sampler_type, bool(interpolated)))
        if factory is None:
            raise NotImplementedError("%s not implemented yet" % self.sampler_type)
        self.sampler = factory(camera, self)
        assert self.sampler is not None

    def finalize_image(self, camera, image):
        image = image.reshape(*camera.resolution, 4)
        if not self.transfer_function.grey_opacity:
            # the non-grey composite never touches alpha (render_source.py:535-536)
            try:
                from ..device import parallel_assign
                parallel_assign(image[:, :, 3], 1)
            except Exception:   # noqa: BLE001  (host-only use of the mirror)
                image[:, :, 3] = 1
        return image

    def __repr__(self):
        return "<Volume Source>:%s transfer_function:%s" % (self.data_source, self._transfer_function)


class KDTreeVolumeSource(VolumeSource):
    volume_method = "KDTree"

    def _get_volume(self):
        if self._volume is None:
            self._volume = AMRKDTree(self.data_source.ds, data_source=self.data_source)
        return self._volume

    def render(self, camera, zbuffer=None):
        self._ensure_volume_fields()
        self.zbuffer = zbuffer
        self.set_sampler(camera)
        assert self.sampler is not None
        if self.check_nans:
            for brick in self.volume.bricks:
                for data in brick.my_data:
                    if np.any(np.isnan(data)):
                        raise RuntimeError
        if self.volume.comm_size > 1:
            # one process per GPU: this rank casts the leaves of its kd subtree; the partial images are
            # composited in view order on the devices (replaces reduce_tree_images)
            from ..parallel import BinarySwap
            nx, ny = camera.resolution
            if len(self.volume.tree.traverse()) > 0:
                bricks = self.volume.device_bricks(self.sampler._ctx, self.sampler._storage)
                self.sampler.cast_bricks(bricks, self.volume.view_order(camera.lens.viewpoint, camera))
            # one compositor per (image size, partial type): a PeerSwap's IPC-shared slab outlives the per-render
            # samplers.  make_compositor falls back to the NCCL route for > 8 ranks, several hosts or no peer access.
            from ..parallel import PeerSwap, make_compositor
            n_fields = len(self.volume.fields)
            fmt = self.sampler.partial_format(n_fields)
            key = (nx, ny, fmt, self.volume.comm_size)
            cache = self.__dict__.setdefault("_swaps", {})
            if key not in cache:
                cache[key] = make_compositor(self.sampler._ctx, self.sampler._frame, self.volume.kd_path(), camera.lens.viewpoint,
                                             self.volume.comm_rank, self.volume.comm_size, fmt=fmt)[0]
            swap = cache[key]
            if isinstance(swap, PeerSwap):
                swap.set_view(self.volume.kd_path(), camera.lens.viewpoint)
            else:
                swap = BinarySwap(self.sampler._ctx, nx, ny, self.volume.kd_path(), camera.lens.viewpoint,
                                  self.volume.comm_rank, self.volume.comm_size)
            self.sampler.flush_parallel(swap, broadcast=True)
        elif self.batched:
            bricks = self.volume.device_bricks(self.sampler._ctx, self.sampler._storage)
            self.sampler.cast_bricks(bricks, self.volume.view_order(camera.lens.viewpoint, camera))
        else:
            for brick in self.volume.traverse(camera.lens.viewpoint):
                self.sampler(brick, num_threads=self.num_threads)
        self.current_image = self.finalize_image(camera, self.sampler.aimage)
        if zbuffer is None:
            self.zbuffer = ZBuffer(self.current_image, np.full(self.current_image.shape[:2], np.inf))
        return self.current_image

    def finalize_image(self, camera, image):
        if self._volume is not None:
            image = self.volume.reduce_tree_images(image, camera.lens.viewpoint, use_opacity=self.transfer_function.grey_opacity)
        return super().finalize_image(camera, image)


class _OctreeVolume(OctreeRayTracing):
    """octree_raytracing.py:9-98 over an ``OctreeAllData``: the tree from the cell centres / levels, and
    ``set_fields`` leaving the vertex values in ``self.data`` as (8, n_cells), x-major within a cell."""

    def __init__(self, data_source, context=None):
        ds = data_source.ds
        xyz = np.stack([data_source[k] for k in "xyz"], axis=-1)
        super().__init__(xyz, data_source["grid_level"], ds.min_level, ds.max_level, context=context)
        self.data_source = data_source

    def set_fields(self, fields, log_fields, no_ghost, force=False):
        if no_ghost:
            raise NotImplementedError("Ghost zones are required with Octree datasets")
        if len(fields) != 1:
            raise ValueError("Can only set one fields at a time. This is likely a bug, and should be reported.")
        field = self.data_source._determine_fields(fields)[0]
        vertex_data = self.data_source.vertex_centered_data(field)
        if log_fields[0]:
            vertex_data = np.log10(vertex_data)
        # "swap axis 0 and 2" then reshape is the reference's layout fix (octree_raytracing.py:93-98): the sampler
        # reads corner (x, y, z) at x*4 + y*2 + z
        self.data = vertex_data.swapaxes(0, 2).reshape(8, -1)


class OctreeVolumeSource(VolumeSource):
    """render_source.py:619-693: ghost zones always on, one PartitionedGrid holding 14 doubles per cell, one
    ``sampler(pg, oct=tree)`` call."""
    volume_method = "Octree"

    def __init__(self, *args, **kwa):
        super().__init__(*args, **kwa)
        self.use_ghost_zones = True

    def _get_volume(self):
        if self._volume is None:
            self._volume = _OctreeVolume(self.data_source)
        return self._volume

    def render(self, camera, zbuffer=None):
        self._ensure_volume_fields()
        self.zbuffer = zbuffer
        self.set_sampler(camera)
        if self.sampler is None:
            raise RuntimeError("No sampler set. This is likely a bug as it should never happen.")
        # the 14 doubles per cell depend only on the volume's field data and the cell geometry: assemble (and upload)
        # them once per set_fields, not once per frame
        token = (id(self.volume.data), id(self.volume.octree))
        if getattr(self, "_cells_pg", None) is None or self._cells_pg[0] != token:
            data = self.data_source
            dx = data["dx"][:, None]
            xyz = np.stack([data[k] for k in "xyz"], axis=-1)
            LE, RE = xyz - dx / 2, xyz + dx / 2
            dt = np.stack(list(self.volume.data) + [*LE.T, *RE.T], axis=-1).reshape(1, len(dx), 14, 1)
            mask = np.full(dt.shape[1:], 1, dtype=np.uint8)
            dims = np.array([1, 1, 1], dtype="int64")
            pg = PartitionedGrid(0, dt, mask, LE.flatten(), RE.flatten(), dims, n_fields=1)
            pg._resident_token = token
            self._cells_pg = (token, pg, self.volume.data)   # keeps volume.data alive: its id stays unique
        pg = self._cells_pg[1]
        self.sampler(pg, oct=self.volume.octree)
        self.current_image = self.finalize_image(camera, self.sampler.aimage)
        if zbuffer is None:
            self.zbuffer = ZBuffer(self.current_image, np.full(self.current_image.shape[:2], np.inf))
        return self.current_image


# ---------------------------------------------------------------------------------------------------------
# opaque sources (render_source.py:989-1591): painted into the ZBuffer before the volume is cast
# ---------------------------------------------------------------------------------------------------------
def _fresh_zbuffer(camera, zbuffer):
    if zbuffer is None:
        empty = camera.lens.new_image(camera)
        empty[:] = 0.0
        return ZBuffer(empty, np.full(empty.shape[:2], np.inf, dtype="float64"))
    return zbuffer


class PointSource(OpaqueSource):
    """Depth-tested points (render_source.py:989-1103): ``positions`` (N,3), ``colors`` (N,4) RGBA, ``radii`` pixels."""
    data_source = None

    def __init__(self, positions, colors=None, color_stride=1, radii=None):
        super().__init__()
        positions = np.asarray(positions, dtype="float64")
        assert positions.ndim == 2 and positions.shape[1] == 3
        if colors is not None:
            colors = np.asarray(colors, dtype="float64")
            assert colors.ndim == 2 and colors.shape[1] == 4
            assert colors.shape[0] == positions.shape[0]
        if radii is None or np.ndim(radii) == 0:
            radii = (0 if radii is None else radii) * np.ones(positions.shape[0], dtype="int64")
        else:
            radii = np.asarray(radii, dtype="int64")
            assert radii.ndim == 1 and radii.shape[0] == positions.shape[0]
        self.positions = positions
        self.colors = np.ones((len(positions), 4)) if colors is None else colors
        self.color_stride = color_stride
        self.radii = radii

    def _validate(self):
        pass

    def render(self, camera, zbuffer=None):
        from ..zraster import zpoints
        zbuffer = _fresh_zbuffer(camera, zbuffer)
        camera.lens.setup_box_properties(camera)
        px, py, dz = camera.lens.project_to_plane(camera, self.positions)
        zpoints(zbuffer.rgba, zbuffer.z, px, py, dz, self.colors, self.radii, self.color_stride)
        self.zbuffer = zbuffer
        return zbuffer

    def __repr__(self):
        return "<Point Source>"


class LineSource(OpaqueSource):
    """Depth-tested line segments (render_source.py:1106-1254): ``positions`` (N,2,3) end points, ``colors`` (N,4)."""
    data_source = None

    def __init__(self, positions, colors=None, color_stride=1):
        super().__init__()
        positions = np.asarray(positions, dtype="float64")
        assert positions.ndim == 3 and positions.shape[1] == 2 and positions.shape[2] == 3
        if colors is not None:
            colors = np.asarray(colors, dtype="float64")
            assert colors.ndim == 2 and colors.shape[1] == 4
        n = positions.shape[0]
        self.positions = positions.reshape((2 * n, 3))
        # the reference sizes the default colours by the RESHAPED positions (render_source.py:1176-1181)
        self.colors = np.ones((len(self.positions), 4)) if colors is None else colors
        self.color_stride = color_stride

    def _validate(self):
        pass

    def render(self, camera, zbuffer=None):
        from ..zraster import zlines
        zbuffer = _fresh_zbuffer(camera, zbuffer)
        camera.lens.setup_box_properties(camera)
        px, py, dz = camera.lens.project_to_plane(camera, self.positions)
        zlines(zbuffer.rgba, zbuffer.z, px.astype("int64"), py.astype("int64"), dz, self.colors.astype("float64"), self.color_stride)
        self.zbuffer = zbuffer
        return zbuffer

    def __repr__(self):
        return "<Line Source>"


_BOX_EDGES = [0, 1, 1, 2, 2, 3, 3, 0] + [4, 5, 5, 6, 6, 7, 7, 4] + [0, 4, 1, 5, 2, 6, 3, 7]


def _box_vertices(corners):
    """corners (8,3,nboxes) -> (12*nboxes, 2, 3) edge end points, the reference's ordering (render_source.py:1302-1309)."""
    nb = corners.shape[2]
    vertices = np.empty([nb * 24, 3])
    for i in range(3):
        vertices[:, i] = corners[_BOX_EDGES, i, ...].ravel(order="F")
    return vertices.reshape((nb * 12, 2, 3))


class BoxSource(LineSource):
    """The 12 edges of a box (render_source.py:1257-1316)."""

    def __init__(self, left_edge, right_edge, color=None):
        left_edge, right_edge = np.asarray(left_edge, dtype="float64"), np.asarray(right_edge, dtype="float64")
        assert left_edge.shape == (3,) and right_edge.shape == (3,)
        color = np.array([[1.0, 1.0, 1.0, 1.0]]) if color is None else np.atleast_2d(np.asarray(color, dtype="float64"))
        assert color.shape == (1, 4)
        from .utils import get_corners
        corners = get_corners(left_edge.copy(), right_edge.copy())
        super().__init__(_box_vertices(corners[:, :, None]), color, color_stride=24)


class GridSource(LineSource):
    """The edges of every grid of a data source, coloured by level (render_source.py:1318-1430)."""

    def __init__(self, data_source, alpha=0.3, cmap=None, min_level=None, max_level=None):
        from . import colormaps
        self.data_source = data_source_or_all(data_source)
        corners, levels = [], []
        for block, _mask in self.data_source.blocks:
            le, re = block.LeftEdge, block.RightEdge
            corners.append(np.array([[le[0], le[1], le[2]], [re[0], le[1], le[2]], [re[0], re[1], le[2]], [le[0], re[1], le[2]],
                                     [le[0], le[1], re[2]], [re[0], le[1], re[2]], [re[0], re[1], re[2]], [le[0], re[1], re[2]]],
                                    dtype="float64"))
            levels.append(block.Level)
        corners = np.dstack(corners)
        levels = np.array(levels)
        if max_level is not None:
            subset = levels <= max_level
            levels, corners = levels[subset], corners[:, :, subset]
        if min_level is not None:
            subset = levels >= min_level
            levels, corners = levels[subset], corners[:, :, subset]
        top = max(float(self.data_source.ds.index.max_level), 1e-300)
        # apply_colormap quantises to 8 bits (image_writer.py); the map itself is an input to both sides
        rgba = colormaps.get(cmap or "arbre")(np.clip(levels / top, 0.0, 1.0))
        colors = np.floor(np.asarray(rgba, dtype="float64") * 255.0) / 255.0
        colors[:, 3] = alpha
        LineSource.__init__(self, _box_vertices(corners), colors, color_stride=24)


class CoordinateVectorSource(OpaqueSource):
    """Small x/y/z axes triad in the image corner (render_source.py:1432-1593)."""
    data_source = None

    def __init__(self, colors=None, alpha=1.0, *, thickness=1):
        super().__init__()
        if colors is None:
            colors = np.zeros((3, 4))
            colors[0, 0] = colors[1, 1] = colors[2, 2] = 1.0
            colors[:, 3] = alpha
        self.colors = np.asarray(colors, dtype="float64")
        self.thick = thickness

    def _validate(self):
        pass

    def render(self, camera, zbuffer=None):
        from ..zraster import zlines
        camera.lens.setup_box_properties(camera)
        positions = np.zeros([6, 3])
        positions[:] = camera.focus
        for i in range(3):
            positions[2 * i + 1, i] += camera.width[i] / 16.0
        px, py, dz = camera.lens.project_to_plane(camera, positions)
        dpx, dpy = px[1::2] - px[::2], py[1::2] - py[::2]
        lpx = camera.resolution[0] / 8
        lpy = camera.resolution[1] - camera.resolution[1] / 8
        px, py = px.astype("float64"), py.astype("float64")
        px[::2], px[1::2] = lpx, lpx + dpx
        py[::2], py[1::2] = lpy, lpy + dpy
        dz = np.zeros_like(dz)
        zbuffer = _fresh_zbuffer(camera, zbuffer)
        zlines(zbuffer.rgba, zbuffer.z, px.astype("int64"), py.astype("int64"), dz, self.colors.astype("float64"), thick=self.thick)
        self.zbuffer = zbuffer
        return zbuffer

    def __repr__(self):
        return "<Coordinate Vector Source>"
