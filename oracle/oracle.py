"""ctypes front-end of the CPU oracle (``ytray_oracle.c``).

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by ``yt_b200``.

The classes mirror the reference boundary so that a test can drive the compiled reference
(``oracle/_ref``), this oracle and the CUDA samplers with the very same arguments:

  * ``PartitionedGrid(parent_grid_id, data, mask, left_edge, right_edge, dims)``
    -- yt/utilities/lib/partitioned_grid.pyx:28-63
  * ``ProjectionSampler / InterpolatedProjectionSampler / VolumeRenderSampler``
    -- yt/utilities/lib/image_samplers.pyx:59-130, 327-343, 346-402, 405-487

``load_ref()`` returns the compiled reference's classes when ``oracle/_ref`` is present.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "_build", "libytray_oracle.so")
_SRC = os.path.join(HERE, "ytray_oracle.c")
_lib = None

PROJ, IPROJ, VR, LIGHT = 0, 1, 2, 3
LENS_PP, LENS_PERSPECTIVE, LENS_GENERIC = 0, 1, 2


def build(force=False):
    """gcc -O2 -ffp-contract=off -fopenmp ytray_oracle.c -> oracle/_build/libytray_oracle.so"""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


class _TFDesc(C.Structure):
    _fields_ = [("n_fits", C.c_int), ("nbins", C.c_int * 6),
                ("values", C.POINTER(C.c_double) * 6), ("bounds", (C.c_double * 2) * 6),
                ("field_ids", C.c_int * 6), ("weight_field_ids", C.c_int * 6),
                ("weight_table_ids", C.c_int * 6), ("field_table_ids", C.c_int * 6),
                ("grey_opacity", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.orc_cast_brick.restype = C.c_int64
        _lib.orc_cast_brick.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
            C.c_int, C.POINTER(dp), C.POINTER(C.c_uint8), dp, dp, C.POINTER(C.c_int64),
            C.POINTER(_TFDesc), C.c_int, C.c_int, dp, dp]
        _lib.orc_max_threads.restype = C.c_int
        i64p = C.POINTER(C.c_int64)
        _lib.orc_ghost_zone_interpolate.restype = None
        _lib.orc_ghost_zone_interpolate.argtypes = [C.c_int, dp, i64p, dp, dp, i64p, dp]
        i32p = C.POINTER(C.c_int32)
        _lib.orc_octree_create.restype = C.c_void_p
        _lib.orc_octree_create.argtypes = [C.c_int, dp, dp]
        _lib.orc_octree_destroy.restype = None
        _lib.orc_octree_destroy.argtypes = [C.c_void_p]
        _lib.orc_octree_add_nodes.restype = C.c_int
        _lib.orc_octree_add_nodes.argtypes = [C.c_void_p, C.c_int64, i32p, i32p, i32p]
        _lib.orc_octree_node_count.restype = C.c_int64
        _lib.orc_octree_node_count.argtypes = [C.c_void_p]
        _lib.orc_octree_cast_ray.restype = C.c_int64
        _lib.orc_octree_cast_ray.argtypes = [C.c_void_p, dp, dp, C.c_int64, i32p, dp]
        _lib.orc_cast_octree.restype = C.c_int64
        _lib.orc_cast_octree.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_void_p,
                                         C.POINTER(_TFDesc), C.c_int, C.c_int]
    return _lib


def max_threads():
    return lib().orc_max_threads()


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class PartitionedGrid:
    def __init__(self, parent_grid_id, data, mask, left_edge, right_edge, dims, n_fields=-1):
        self.parent_grid_id = parent_grid_id
        self.my_data = [np.ascontiguousarray(d, dtype="float64") for d in data]
        self.source_mask = np.ascontiguousarray(mask, dtype="uint8")
        self.LeftEdge = np.ascontiguousarray(left_edge, dtype="float64")
        self.RightEdge = np.ascontiguousarray(right_edge, dtype="float64")
        self.dims = np.ascontiguousarray(dims, dtype="int64")
        self.n_fields = len(self.my_data) if n_fields == -1 else n_fields


class ImageSampler:
    kind = None

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method,
                 *args, **kwargs):
        self.volume_method = volume_method
        self.lens_type = kwargs.pop("lens_type", None)
        camera_data = kwargs.pop("camera_data", None)
        zbuffer = kwargs.pop("zbuffer", None)
        if not (isinstance(image, np.ndarray) and image.dtype == np.float64 and image.ndim == 3
                and image.flags.c_contiguous):
            raise ValueError("image must be a C-contiguous float64 (nx, ny, 4) array")
        self.aimage = image
        nx, ny = image.shape[:2]
        self.vp_pos = np.ascontiguousarray(vp_pos, dtype="float64")
        self.vp_dir = np.ascontiguousarray(vp_dir, dtype="float64")
        if self.lens_type == "plane-parallel":
            self._lens = LENS_PP
        else:
            if self.vp_pos.shape[:2] != (nx, ny) or self.vp_dir.shape[:2] != (nx, ny):
                raise RuntimeError("Bad lens shape / direction for %s" % self.lens_type)
            self._lens = (LENS_PERSPECTIVE if (camera_data is not None and self.lens_type == "perspective")
                          else LENS_GENERIC)
        self.camera_data = None if camera_data is None else np.ascontiguousarray(camera_data, dtype="float64")
        self.azbuffer = (np.ones((nx, ny), "float64") if zbuffer is None
                         else np.ascontiguousarray(zbuffer, dtype="float64"))
        self.acenter = np.ascontiguousarray(center, dtype="float64")
        self.ax_vec = np.ascontiguousarray(x_vec, dtype="float64")
        self.ay_vec = np.ascontiguousarray(y_vec, dtype="float64")
        self.bounds = np.array([float(b) for b in bounds], dtype="float64")
        self.width = np.ascontiguousarray(width, dtype="float64")
        self.n_samples = 10
        self._tf = None
        self.cells = 0

    def cast_through_octree(self, pg, oct, num_threads=0):
        """image_samplers.pyx:202-299.  `pg` holds one field of shape (n_cells, 14, 1): 8 vertex values, left edge,
        right edge per cell (render_source.py:669-680); the rays are the per-pixel (nx, ny, 3) arrays."""
        nx, ny = self.aimage.shape[:2]
        if self.vp_pos.shape != (nx, ny, 3) or self.vp_dir.shape != (nx, ny, 3):
            raise ValueError("cast_through_octree reads per-pixel vp_pos / vp_dir of shape (nx, ny, 3)")
        cells = np.ascontiguousarray(pg.my_data[0], dtype="float64").reshape(-1, 14)
        n = lib().orc_cast_octree(self.kind, nx, ny, _dptr(self.vp_pos), _dptr(self.vp_dir), _dptr(self.aimage),
                                  _dptr(cells), oct._h, C.byref(self._tf) if self._tf is not None else None,
                                  int(self.n_samples), int(num_threads))
        if n < 0:
            raise RuntimeError("oracle error %d" % n)
        self.cells += n
        return None

    def __call__(self, pg, num_threads=0, oct=None):
        if self.volume_method == "Octree":
            return self.cast_through_octree(pg, oct, num_threads=num_threads)
        if self.volume_method != "KDTree":
            raise NotImplementedError(
                'Volume rendering has not been implemented for method: "%s"' % self.volume_method)
        nf = pg.n_fields
        arr = (C.POINTER(C.c_double) * nf)(*[_dptr(d) for d in pg.my_data[:nf]])
        nx, ny = self.aimage.shape[:2]
        n = lib().orc_cast_brick(
            self.kind, self._lens, nx, ny, _dptr(self.vp_pos), _dptr(self.vp_dir), _dptr(self.acenter),
            _dptr(self.bounds), _dptr(self.ax_vec), _dptr(self.ay_vec), _dptr(self.width),
            _dptr(self.camera_data), _dptr(self.aimage), _dptr(self.azbuffer),
            nf, arr, pg.source_mask.ctypes.data_as(C.POINTER(C.c_uint8)),
            _dptr(pg.LeftEdge), _dptr(pg.RightEdge), pg.dims.ctypes.data_as(C.POINTER(C.c_int64)),
            C.byref(self._tf) if self._tf is not None else None, int(self.n_samples), int(num_threads),
            _dptr(getattr(self, "light_dir", None)), _dptr(getattr(self, "light_rgba", None)))
        if n < 0:
            raise RuntimeError("oracle error %d" % n)
        self.cells += n
        return 0


class ProjectionSampler(ImageSampler):
    kind = PROJ


class InterpolatedProjectionSampler(ImageSampler):
    kind = IPROJ

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method,
                 n_samples=10, **kwargs):
        super().__init__(vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, **kwargs)
        self.n_samples = n_samples


class VolumeRenderSampler(ImageSampler):
    kind = VR

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method,
                 tf_obj, n_samples=10, **kwargs):
        super().__init__(vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, **kwargs)
        self.n_samples = n_samples
        d = _TFDesc()
        d.n_fits = tf_obj.n_field_tables
        assert d.n_fits <= 6
        self._keep = []
        for i in range(d.n_fits):
            y = np.ascontiguousarray(tf_obj.tables[i].y, dtype="float64")
            self._keep.append(y)
            d.nbins[i] = y.shape[0]
            d.values[i] = _dptr(y)
            d.bounds[i][0] = float(tf_obj.tables[i].x_bounds[0])
            d.bounds[i][1] = float(tf_obj.tables[i].x_bounds[1])
            d.field_ids[i] = tf_obj.field_ids[i]
            d.weight_field_ids[i] = tf_obj.weight_field_ids[i]
            d.weight_table_ids[i] = tf_obj.weight_table_ids[i]
        for i in range(6):
            d.field_table_ids[i] = tf_obj.field_table_ids[i]
        d.grey_opacity = int(getattr(tf_obj, "grey_opacity", 0))
        self._tf = d


class LightSourceRenderSampler(VolumeRenderSampler):
    """image_samplers.pyx:496-600: the volume render lit by one directional light (gradient shading)."""
    kind = LIGHT

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method,
                 tf_obj, n_samples=10, light_dir=(1., 1., 1.), light_rgba=(1., 1., 1., 1.), **kwargs):
        super().__init__(vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, tf_obj, n_samples, **kwargs)
        ld = np.array(light_dir, dtype="float64")
        ld /= np.sqrt(ld[0] * ld[0] + ld[1] * ld[1] + ld[2] * ld[2])     # image_samplers.pyx:526-528
        self.light_dir = np.ascontiguousarray(ld)
        self.light_rgba = np.ascontiguousarray(np.array(light_rgba, dtype="float64"))


class _OctreeRayTracing:
    """_octree_raytracing.pyx:15-34 over the restated Octree<int> (_octree_raytracing.hpp)."""

    def __init__(self, LE, RE, depth):
        self._LE = np.ascontiguousarray(LE, dtype="float64")
        self._RE = np.ascontiguousarray(RE, dtype="float64")
        self.depth = int(depth)
        self._h = C.c_void_p(lib().orc_octree_create(self.depth, _dptr(self._LE), _dptr(self._RE)))

    def add_nodes(self, ipos, lvl, key):
        ipos = np.ascontiguousarray(ipos, dtype="int32")
        lvl = np.ascontiguousarray(lvl, dtype="int32")
        key = np.ascontiguousarray(key, dtype="int32")
        i32p = C.POINTER(C.c_int32)
        rc = lib().orc_octree_add_nodes(self._h, len(key), ipos.ctypes.data_as(i32p), lvl.ctypes.data_as(i32p),
                                        key.ctypes.data_as(i32p))
        if rc:
            raise ValueError("level beyond the octree depth")

    def cast_ray(self, o, d, cap=4096):
        """(keys, t[:, 2]) of one ray: Octree::cast_ray(double*, double*, keyList, tList)."""
        o = np.ascontiguousarray(o, dtype="float64"); d = np.ascontiguousarray(d, dtype="float64")
        keys = np.empty(cap, "int32"); t = np.empty(2 * cap, "float64")
        n = lib().orc_octree_cast_ray(self._h, _dptr(o), _dptr(d), cap, keys.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(t))
        assert n <= cap
        return keys[:n].copy(), t[:2 * n].reshape(n, 2).copy()

    def __del__(self):
        try:
            lib().orc_octree_destroy(self._h)
        except Exception:
            pass


def vertex_centre(cells):
    """grid.get_vertex_centered_data(no_ghost=True) of one field: grid_patch.py:284-296 ->
    ghost_zone_interpolate(1, cells, [0.5]*3, out, [0]*3) (interpolators.pyx:177-227)."""
    cells = np.ascontiguousarray(cells, dtype="float64")
    out = np.zeros(tuple(s + 1 for s in cells.shape), dtype="float64")
    ish = np.array(cells.shape, dtype="int64")
    osh = np.array(out.shape, dtype="int64")
    il, ol = np.full(3, 0.5), np.zeros(3)
    i64p = C.POINTER(C.c_int64)
    lib().orc_ghost_zone_interpolate(1, _dptr(cells), ish.ctypes.data_as(i64p), _dptr(il), _dptr(out),
                                     osh.ctypes.data_as(i64p), _dptr(ol))
    return out


def brick_data_from_cells(cells, li, ri, take_log):
    """The data array of one kd leaf as AMRKDTree.get_brick_data builds it (amr_kdtree.py:343-365):
    vertex-centre the whole grid, log10 with values <= 0 -> NaN, slice [li:ri+1]."""
    v = vertex_centre(cells)
    if take_log:
        v[v <= 0] = np.nan
        with np.errstate(invalid="ignore"):
            v = np.log10(v)
    return v[li[0]:ri[0] + 1, li[1]:ri[1] + 1, li[2]:ri[2] + 1].copy()


def ref_available():
    import sysconfig
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.exists(os.path.join(HERE, "_ref", "yt", "utilities", "lib", "image_samplers" + suffix))


def load_ref():
    """Import the UNMODIFIED reference samplers compiled into oracle/_ref (see build_ref.py)."""
    if not ref_available():
        raise ImportError("oracle/_ref is not built (python oracle/build_ref.py)")
    p = os.path.join(HERE, "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    from yt.utilities.lib import image_samplers, partitioned_grid  # noqa
    return image_samplers, partitioned_grid
