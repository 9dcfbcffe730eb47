"""zlines / zpoints -- the opaque-source painters of yt/utilities/lib/misc_utilities.pyx:244-404, same call
signatures, executed by libytb200 (``csrc/zraster.cpp``).  They paint depth-tested line segments / discs into the
host ``(nx, ny, 4)`` image and ``(nx, ny)`` z-buffer of a :class:`~yt_b200.volume_rendering.zbuffer_array.ZBuffer`;
the volume cast then starts from those colours and stops at those depths."""
import numpy as np

from . import _lib as L


def _check(image, zbuffer):
    if not (isinstance(image, np.ndarray) and image.dtype == np.float64 and image.ndim == 3 and image.shape[2] == 4
            and image.flags.c_contiguous):
        raise ValueError("image must be a C-contiguous float64 (nx, ny, 4) array")
    if not (isinstance(zbuffer, np.ndarray) and zbuffer.dtype == np.float64 and zbuffer.shape == image.shape[:2]
            and zbuffer.flags.c_contiguous):
        raise ValueError("zbuffer must be a C-contiguous float64 (nx, ny) array")


def zlines(image, zbuffer, xs, ys, zs, colors, points_per_color=1, thick=1, flip=0, crop=0):
    _check(image, zbuffer)
    xs = np.ascontiguousarray(xs, dtype=np.int64)
    ys = np.ascontiguousarray(ys, dtype=np.int64)
    zs = np.ascontiguousarray(zs, dtype=np.float64)
    colors = np.ascontiguousarray(colors, dtype=np.float64)
    if not (xs.shape == ys.shape == zs.shape and xs.ndim == 1) or colors.ndim != 2 or colors.shape[1] != 4:
        raise ValueError("xs, ys, zs must be 1-D arrays of one length and colors (n, 4)")
    rc = L.load().ytb200_zlines(image.ctypes.data_as(L.dp), zbuffer.ctypes.data_as(L.dp), image.shape[0], image.shape[1],
                                xs.ctypes.data_as(L.i64p), ys.ctypes.data_as(L.i64p), zs.ctypes.data_as(L.dp), xs.shape[0],
                                colors.ctypes.data_as(L.dp), colors.shape[0], int(points_per_color), int(thick), int(flip), int(crop))
    if rc != 0:
        raise ValueError("zlines: bad arguments (odd number of points, or fewer colours than segments)")


def zpoints(image, zbuffer, xs, ys, zs, colors, radii, points_per_color=1, thick=1, flip=0):
    _check(image, zbuffer)
    xs = np.ascontiguousarray(xs, dtype=np.int64)
    ys = np.ascontiguousarray(ys, dtype=np.int64)
    zs = np.ascontiguousarray(zs, dtype=np.float64)
    colors = np.ascontiguousarray(colors, dtype=np.float64)
    radii = np.ascontiguousarray(radii, dtype=np.int64)
    order = np.ascontiguousarray(np.argsort(zs), dtype=np.int64)   # misc_utilities.pyx:363
    rc = L.load().ytb200_zpoints(image.ctypes.data_as(L.dp), zbuffer.ctypes.data_as(L.dp), image.shape[0], image.shape[1],
                                 xs.ctypes.data_as(L.i64p), ys.ctypes.data_as(L.i64p), zs.ctypes.data_as(L.dp), xs.shape[0],
                                 order.ctypes.data_as(L.i64p), colors.ctypes.data_as(L.dp), colors.shape[0],
                                 radii.ctypes.data_as(L.i64p), int(points_per_color), int(flip))
    if rc != 0:
        raise ValueError("zpoints: bad arguments")
