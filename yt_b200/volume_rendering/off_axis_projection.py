"""``off_axis_projection`` for grid datasets -- yt/visualization/volume_rendering/off_axis_projection.py:21-547.

The grid branch (506-526): every grid of the data source becomes one cell-centred PartitionedGrid
with its data pre-multiplied by the (selector & child) mask and is cast through ProjectionSampler;
here all grids go to the device as one brick table and one launch.
"""
import numpy as np

from ..partitioned_grid import PartitionedGrid
from .image_array import ImageArray
from .render_source import KDTreeVolumeSource
from .scene import Scene
from .transfer_functions import ProjectionTransferFunction
from .utils import data_source_or_all


def off_axis_projection(data_source, center, normal_vector, width, resolution, item, weight=None, volume=None,
                        no_ghost=False, interpolated=False, north_vector=None, depth=None, num_threads=1, method="integrate"):
    if method not in ("integrate", "sum"):
        raise NotImplementedError("Only 'integrate' or 'sum' methods are valid for off-axis-projections")
    if interpolated:
        raise NotImplementedError("Only interpolated=False methods are currently implemented for off-axis-projections")
    data_source = data_source_or_all(data_source)
    ds = data_source.ds
    item = data_source._determine_fields([item])[0]
    center = np.array(center, dtype="float64")
    sc = Scene()
    vol = KDTreeVolumeSource(data_source, item)
    vol.num_threads = num_threads
    fields = [item]
    wkey = None
    if weight is not None:
        wkey = data_source._determine_fields([weight])[0]
        fields = [("index", "temp_weightfield"), wkey]
    vol.set_transfer_function(ProjectionTransferFunction())
    camera = sc.add_camera(data_source)
    if np.ndim(width) == 0:
        width = np.array([width] * 3, dtype="float64")
    width = np.array(width, dtype="float64")
    if depth is not None:
        width[2] = depth
    camera.set_width(width)
    if np.ndim(resolution) == 0:
        resolution = [resolution] * 2
    camera.resolution = resolution
    normal = np.array(normal_vector, dtype="float64")
    normal = normal / np.linalg.norm(normal)
    camera.position = center - width[2] * normal
    camera.focus = center
    if north_vector is None:
        vecs = np.identity(3)
        ax = np.cross(vecs, normal).sum(axis=1).argmax()
        east_vector = np.cross(vecs[ax, :], normal).ravel()
        north = np.cross(normal, east_vector).ravel()
    else:
        north = np.array(north_vector, dtype="float64")
        north = north / np.linalg.norm(north)
    camera.switch_orientation(normal, north)
    sc.add_source(vol)
    vol.set_sampler(camera, interpolated=False)
    bricks = []
    for grid, mask in data_source.blocks:
        data = []
        for f in fields:
            if f == ("index", "temp_weightfield"):
                gd = grid[item].astype("float64") * grid[wkey]
            else:
                gd = grid[f]
            data.append(np.ascontiguousarray(gd * mask, dtype="float64"))
        bricks.append(PartitionedGrid(grid.id, data, mask.astype("uint8"), grid.LeftEdge, grid.RightEdge,
                                      grid.ActiveDimensions.astype("int64")))
    for pg in bricks:
        vol.sampler(pg, num_threads=num_threads)
    image = vol.sampler.aimage.reshape(*camera.resolution, 4)
    image[:, :, 3] = 1
    image = ImageArray(image, info={"imtype": "rendering"})
    if method == "integrate":
        if weight is None:
            image *= width[2]
        else:
            zero = image[:, :, 1] == 0
            nz = np.logical_not(zero)
            image[:, :, 0][nz] /= image[:, :, 1][nz]
            image[zero] = 0
    return image[:, :, 0]
