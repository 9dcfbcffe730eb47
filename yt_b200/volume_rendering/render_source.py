"""Render sources -- yt/visualization/volume_rendering/render_source.py.

``VolumeSource`` / ``KDTreeVolumeSource`` (209-616) with the reference's properties, setters and
``render``/``finalize_image`` contract.  The brick loop of ``render`` (594-597,
``for brick in volume.traverse(viewpoint): sampler(brick)``) is handed to the device as ONE launch:
all kd-tree leaves are kept resident as a brick table and only the view order is recomputed per
frame (``batched=True``, the default); ``batched=False`` drives the sampler brick by brick exactly
like the reference does.
"""
import abc
import os

import numpy as np

from ..amr_kdtree import AMRKDTree
from .transfer_function_helper import TransferFunctionHelper
from .transfer_functions import ColorTransferFunction, ProjectionTransferFunction, TransferFunction
from .utils import data_source_or_all, new_interpolated_projection_sampler, new_projection_sampler, new_volume_render_sampler
from .zbuffer_array import ZBuffer


def invalidate_volume(f):
    def wrapper(*args, **kwargs):
        ret = f(*args, **kwargs)
        obj = args[0]
        if isinstance(obj._transfer_function, ProjectionTransferFunction):
            obj.sampler_type = "projection"
            obj._log_field = False
            obj._use_ghost_zones = False
        del obj.volume
        obj._volume_valid = False
        return ret
    return wrapper


def validate_volume(f):
    def wrapper(*args, **kwargs):
        obj = args[0]
        fields = [obj.field]
        log_fields = [obj.log_field]
        if obj.weight_field is not None:
            fields.append(obj.weight_field)
            log_fields.append(obj.log_field)
        if not obj._volume_valid:
            obj.volume.set_fields(fields, log_fields, no_ghost=(not obj.use_ghost_zones))
        obj._volume_valid = True
        return f(*args, **kwargs)
    return wrapper


class RenderSource(abc.ABC):
    volume_method = None

    def __init__(self):
        self.opaque = False
        self.zbuffer = None

    @abc.abstractmethod
    def render(self, camera, zbuffer=None):
        pass

    @abc.abstractmethod
    def _validate(self):
        pass


class OpaqueSource(RenderSource):
    def __init__(self):
        super().__init__()
        self.opaque = True

    def set_zbuffer(self, zbuffer):
        self.zbuffer = zbuffer


def create_volume_source(data_source, field):
    data_source = data_source_or_all(data_source)
    if hasattr(data_source.ds.index, "grids"):
        return KDTreeVolumeSource(data_source, field)
    raise NotImplementedError("only grid (kd-tree) datasets are on the B200 hot path")


class VolumeSource(RenderSource, abc.ABC):
    _image = None
    data_source = None

    def __init__(self, data_source, field):
        super().__init__()
        self.data_source = data_source_or_all(data_source)
        field = self.data_source._determine_fields(field)[0]
        self.current_image = None
        self.check_nans = False
        self.num_threads = 0
        self.num_samples = 10
        self.sampler_type = "volume-render"
        self.batched = True
        self._volume_valid = False
        self._volume = None
        self._transfer_function = None
        self._field = field
        self._log_field = self.data_source.ds.field_info[field].take_log
        self._use_ghost_zones = False
        self._weight_field = None
        self.tfh = TransferFunctionHelper(self.data_source.pf)
        self.tfh.set_field(self.field)

    # -- transfer function ------------------------------------------------------------------------------
    @property
    def transfer_function(self):
        if self._transfer_function is not None:
            return self._transfer_function
        if self.tfh.tf is not None:
            self._transfer_function = self.tfh.tf
            return self._transfer_function
        self.tfh.set_field(self.field)
        self.tfh.set_log(self.log_field)
        self.tfh.build_transfer_function()
        self.tfh.setup_default()
        self._transfer_function = self.tfh.tf
        return self._transfer_function

    @transfer_function.setter
    def transfer_function(self, value):
        self.tfh.tf = None
        valid_types = (TransferFunction, ColorTransferFunction, ProjectionTransferFunction, type(None))
        if not isinstance(value, valid_types):
            raise RuntimeError("transfer_function not a valid type, received object of type %s" % type(value))
        if isinstance(value, ProjectionTransferFunction):
            self.sampler_type = "projection"
            if self._volume is not None:
                self._volume_valid = False
        self._transfer_function = value

    # -- volume ---------------------------------------------------------------------------------------------
    @property
    def volume(self):
        return self._get_volume()

    @volume.setter
    def volume(self, value):
        assert isinstance(value, AMRKDTree)
        self._field = value.fields
        self._log_field = value.log_fields
        self._volume = value
        assert self._volume_valid

    @volume.deleter
    def volume(self):
        self._volume = None

    @property
    def field(self):
        return self._field

    @field.setter
    @invalidate_volume
    def field(self, value):
        field = self.data_source._determine_fields(value)
        if len(field) > 1:
            raise RuntimeError("VolumeSource.field can only be a single field but received multiple fields: %s" % field)
        field = field[0]
        if self._field != field:
            log_field = self.data_source.ds.field_info[field].take_log
            self.tfh.bounds = None
        else:
            log_field = self._log_field
        self._log_field = log_field
        self._field = field
        self.transfer_function = None
        self.tfh.set_field(field)
        self.tfh.set_log(log_field)

    @property
    def log_field(self):
        return self._log_field

    @log_field.setter
    @invalidate_volume
    def log_field(self, value):
        self.transfer_function = None
        self.tfh.set_log(value)
        self._log_field = value

    @property
    def use_ghost_zones(self):
        return self._use_ghost_zones

    @use_ghost_zones.setter
    @invalidate_volume
    def use_ghost_zones(self, value):
        self._use_ghost_zones = value

    @property
    def weight_field(self):
        return self._weight_field

    @weight_field.setter
    @invalidate_volume
    def weight_field(self, value):
        self._weight_field = value

    def set_transfer_function(self, transfer_function):
        self.transfer_function = transfer_function
        return self

    def _validate(self):
        if self.data_source is None:
            raise RuntimeError("Data source not initialized")

    def set_volume(self, volume):
        self.volume = volume
        return self

    def set_field(self, field):
        self.field = field
        return self

    def set_log(self, log_field):
        self.log_field = log_field
        return self

    def set_weight_field(self, weight_field):
        self.weight_field = weight_field
        return self

    def set_use_ghost_zones(self, use_ghost_zones):
        self.use_ghost_zones = use_ghost_zones
        return self

    def set_sampler(self, camera, interpolated=True):
        if self.sampler_type == "volume-render":
            sampler = new_volume_render_sampler(camera, self)
        elif self.sampler_type == "projection" and interpolated:
            sampler = new_interpolated_projection_sampler(camera, self)
        elif self.sampler_type == "projection":
            sampler = new_projection_sampler(camera, self)
        else:
            raise NotImplementedError("%s not implemented yet" % self.sampler_type)
        self.sampler = sampler
        assert self.sampler is not None

    @abc.abstractmethod
    def _get_volume(self):
        pass

    def finalize_image(self, camera, image):
        image = image.reshape(*camera.resolution, 4)
        if not self.transfer_function.grey_opacity:
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

    @validate_volume
    def render(self, camera, zbuffer=None):
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
            if os.environ.get("YTB200_SWAP", "peer") == "nccl":
                swap = BinarySwap(self.sampler._ctx, nx, ny, self.volume.kd_path(), camera.lens.viewpoint,
                                  self.volume.comm_rank, self.volume.comm_size)
            else:
                swap = self.sampler.peer_swap(self.volume.kd_path(), camera.lens.viewpoint,
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
