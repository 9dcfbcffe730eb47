"""Drop-in samplers: the boundary of yt/utilities/lib/image_samplers.pyx, on a B200.

``ProjectionSampler``, ``InterpolatedProjectionSampler`` and ``VolumeRenderSampler`` take the
reference's constructor arguments (image_samplers.pyx:59-130, 347-359, 406-419), are called once
per brick in kd-tree order (render_source.py:594-597) and expose the result as ``.aimage`` --
the *same* ndarray that was passed in, updated in place (image_samplers.pyx:111,192).

A GPU wants the whole brick list in one launch, so ``sampler(brick)`` only records the brick;
the cast happens on the first access of ``.aimage`` (or an explicit :meth:`flush`), in the
order of the calls.  ``cast_bricks(brickset, order)`` is the batched entry the scene mirror uses.

All compute is in libytb200.so (CUDA, sm_100a).  There is no CPU path here.
"""
import ctypes as C

import os

import numpy as np

from . import _lib as L
from . import device as D
from .partitioned_grid import PartitionedGrid, build_brickset

#: module-wide defaults; "f32"/"fast" is the product path, "f64"/"exact" reproduces the
#: reference's fp64 evaluation order bit-for-bit (validation mode).
#: "early_termination": threshold of the early-ray-termination cast (ours; 0 = off, the default): grey-opacity
#: ColorTransferFunction renders stop a ray once what it has composited -- walking the reference's composition order
#: from its far end -- transmits less than this (``ytb200_frame_set_early_termination``).  ``YTB200_ERT`` overrides.
config = {"storage": "f32", "precision": "fast", "early_termination": 0.0}

_STORAGE = {"f32": L.STORE_F32, "f64": L.STORE_F64}
_PRECISION = {"fast": L.PREC_FAST, "exact": L.PREC_EXACT}


def _memview3(a, name):
    """Checks of a ``np.float64_t[:,:,:]`` typed memoryview argument."""
    if not isinstance(a, np.ndarray):
        a = np.asarray(a)
    if a.ndim != 3:
        raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % a.ndim)
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % a.dtype)
    if not a.flags.writeable:
        raise ValueError("buffer source array is read-only")
    return a


def _ndarray(a, ndim, name):
    """Checks of a ``np.ndarray[np.float64_t, ndim=N]`` argument."""
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'float64_t' but got '%s'" % a.dtype)
    return a


class ImageSampler:
    kind = None

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, *args, **kwargs):
        if not isinstance(volume_method, str):
            raise TypeError("Argument 'volume_method' has incorrect type (expected str)")
        vp_pos = _memview3(vp_pos, "vp_pos")
        vp_dir = _memview3(vp_dir, "vp_dir")
        if center is not None:   # FisheyeLens passes None (lens.py:530,559); Cython's typed ndarray arguments accept it
            center = _ndarray(center, 1, "center")
        image = _ndarray(image, 3, "image")
        x_vec = _ndarray(x_vec, 1, "x_vec")
        y_vec = _ndarray(y_vec, 1, "y_vec")
        width = _ndarray(width, 1, "width")
        self.volume_method = volume_method
        camera_data = kwargs.pop("camera_data", None)
        self.camera_data = None if camera_data is None else np.ascontiguousarray(camera_data, dtype=np.float64)
        zbuffer = kwargs.pop("zbuffer", None)
        if zbuffer is None:
            zbuffer = np.ones((image.shape[0], image.shape[1]), "float64")
        self.lens_type = kwargs.pop("lens_type", None)
        self._ctx = kwargs.pop("context", None) or D.get_context()
        self._storage = _STORAGE[kwargs.pop("storage", config["storage"])]
        self._precision = _PRECISION[kwargs.pop("precision", config["precision"])]
        self._ert = float(os.environ.get("YTB200_ERT", kwargs.pop("early_termination", config["early_termination"])) or 0.0)
        nx, ny = image.shape[0], image.shape[1]
        if self.lens_type == "plane-parallel":
            lens = L.LENS_PLANE_PARALLEL
        else:
            if not (vp_pos.shape[0] == vp_dir.shape[0] == nx) or not (vp_pos.shape[1] == vp_dir.shape[1] == ny):
                msg = "Bad lens shape / direction for %s\n" % (self.lens_type)
                msg += "Shapes: (%s - %s - %s) and (%s - %s - %s)" % (
                    vp_pos.shape[0], vp_dir.shape[0], nx, vp_pos.shape[1], vp_dir.shape[1], ny)
                raise RuntimeError(msg)
            if camera_data is not None and self.lens_type == "perspective":
                lens = L.LENS_PERSPECTIVE
            else:
                lens = L.LENS_GENERIC
        if image.shape[2] != 4 or not image.flags.c_contiguous:
            raise ValueError("image must be a C-contiguous (nx, ny, 4) float64 array")
        self.vp_pos, self.vp_dir = vp_pos, vp_dir
        self.aimage_host = image
        self.acenter, self.ax_vec, self.ay_vec = center, x_vec, y_vec
        self.azbuffer = np.ascontiguousarray(zbuffer, dtype=np.float64)
        self.aimage_used = np.zeros((nx, ny), "int64")
        self.amesh_lines = np.zeros((nx, ny), "int64")
        self.bounds = tuple(float(b) for b in bounds)
        self.width = np.array(width, dtype=np.float64)
        self.n_samples = 10
        self._tables = None
        # ---- C-ABI camera struct + device frame ----
        cam = L.Camera()
        cam.lens, cam.nx, cam.ny = lens, nx, ny
        keep = []
        if lens == L.LENS_PLANE_PARALLEL:
            flat_p, flat_d = vp_pos.ravel(), vp_dir.ravel()
            if flat_p.shape[0] < 12 or flat_d.shape[0] < 3:
                raise RuntimeError("Bad lens shape / direction for plane-parallel")
            for i in range(12):
                cam.vp_pos[i] = flat_p[i]
            for i in range(3):
                cam.vp_dir[i] = flat_d[i]
        else:
            p = np.ascontiguousarray(vp_pos[:, :, :3])
            d = np.ascontiguousarray(vp_dir[:, :, :3])
            keep += [p, d]
            cam.vp_pos_px = p.ctypes.data_as(L.dp)
            cam.vp_dir_px = d.ctypes.data_as(L.dp)
        for i in range(3):
            cam.center[i] = 0.0 if center is None else center[i]
            cam.x_vec[i], cam.y_vec[i], cam.width[i] = x_vec[i], y_vec[i], width[i]
        for i in range(4):
            cam.bounds[i] = self.bounds[i]
        if self.camera_data is not None:
            cam.has_camera_data = 1
            cd = self.camera_data.ravel()
            for i in range(15):
                cam.camera_data[i] = cd[i]
        self._frame = D.Frame(self._ctx, cam, keep=tuple(keep))
        if self._ert > 0.0:
            self._frame.set_early_termination(self._ert)
        self._pending = []         # PartitionedGrids recorded by __call__, in call order
        self._batches = []         # (brickset, order) recorded by cast_bricks
        self._dirty = False
        self.last_stats = None

    # -- the reference's per-brick entry (image_samplers.pyx:132-144) ---------------------------------
    def __call__(self, pg, **kwa):
        if self.volume_method == "KDTree":
            return self.cast_through_kdtree(pg, **kwa)
        elif self.volume_method == "Octree":
            return self.cast_through_octree(pg, **kwa)
        else:
            raise NotImplementedError(
                'Volume rendering has not been implemented for method: "%s"' % self.volume_method)

    def cast_through_kdtree(self, pg, num_threads=0):
        if not isinstance(pg, PartitionedGrid):
            raise TypeError("Argument 'pg' has incorrect type (expected PartitionedGrid, got %s)" % type(pg).__name__)
        self._flush_batches_if_needed()
        self._pending.append(pg)
        self._dirty = True
        return 0

    def cast_through_octree(self, pg, oct, num_threads=0):
        """image_samplers.pyx:202-299: every pixel's ray walks the octree front to back and the sampler's per-cell body
        runs on each terminal cell, over ``pg``'s single field of shape (n_cells, 14, 1) -- 8 vertex values, left
        edge, right edge per cell (render_source.py:669-680).  The z-buffer is not consulted (as in the reference).
        Runs at once (one launch) and leaves the result in the caller's image.  ``precision="fast"`` (default) samples
        the volume render in fp32, ``"exact"`` does everything in fp64 in the reference's order; the projection kinds
        are fp64 either way."""
        from .octree_raytracing import _OctreeRayTracing
        if not isinstance(pg, PartitionedGrid):
            raise TypeError("Argument 'pg' has incorrect type (expected PartitionedGrid, got %s)" % type(pg).__name__)
        if not isinstance(oct, _OctreeRayTracing):
            raise TypeError("Argument 'oct' has incorrect type (expected _OctreeRayTracing, got %s)" % type(oct).__name__)
        nx, ny = self.aimage_host.shape[:2]
        if self.vp_pos.shape[:2] != (nx, ny) or self.vp_dir.shape[:2] != (nx, ny) or self.lens_type == "plane-parallel":
            # the reference indexes vp_pos[vi, vj, :] unchecked: with the packed plane-parallel description
            # (lens.py:89-94) it reads out of bounds; here that is an error
            raise ValueError("cast_through_octree needs per-pixel vp_pos / vp_dir of shape (nx, ny, 3)")
        if hasattr(self, "light_dir"):
            raise NotImplementedError("LightSourceRenderSampler has no octree cast here")
        cells = pg.my_data[0]
        if cells.size % 14:
            raise ValueError("octree cell data must hold 14 values per cell")
        self.flush()
        oct.set_cells(cells, token=getattr(pg, "_resident_token", None))
        fr = self._frame
        fr.upload(self.aimage_host, self.azbuffer)
        L.check(L.load().ytb200_frame_cast_octree(fr.handle, int(self.kind), oct.handle,
                                                  self._tables.handle if self._tables is not None else None,
                                                  int(self.n_samples), int(self._precision), None))
        fr.download(self.aimage_host)

    # -- batched entry used by the scene mirror ---------------------------------------------------------
    def cast_bricks(self, brickset, order):
        """Record ``for i in order: sampler(brick_i)`` over an already device-resident brick set."""
        self._flush_pending_if_needed()
        self._batches.append((brickset, np.ascontiguousarray(order, dtype=np.int64)))
        self._dirty = True

    def _flush_pending_if_needed(self):
        if self._pending:
            self.flush()

    def _flush_batches_if_needed(self):
        if self._batches:
            self.flush()

    def flush(self, timed=False):
        """Run everything recorded so far on the device and write the result into the host image."""
        if not self._dirty:
            return
        fr = self._frame
        work = list(self._batches)
        if self._pending:
            bs = build_brickset(self._pending, self._ctx, self._storage)
            work.append((bs, np.arange(len(self._pending), dtype=np.int64)))
        if len(work) == 1 and not timed:
            # one brick list: the single C call the e2e benchmark times.  Page-locked image / z-buffer arrays (the scene
            # mirror allocates them so) are read and written in place by the cast kernel; pageable ones are staged.
            bs, order = work[0]
            fr.render(self.kind, bs, order, self.aimage_host, zbuffer=self.azbuffer, tf=self._tables, n_samples=self.n_samples,
                      precision=self._precision)
        else:
            fr.upload(self.aimage_host, self.azbuffer)
            for bs, order in work:
                self.last_stats = fr.cast(self.kind, bs, order, tf=self._tables, n_samples=self.n_samples,
                                          precision=self._precision, timed=timed)
            fr.download(self.aimage_host)
        self._pending, self._batches, self._dirty = [], [], False

    def partial_format(self, n_fields=1):
        """Format of this sampler's partial images (``ytb200_partial_format``)."""
        from .parallel import partial_format
        return partial_format(self.kind, self._tables, n_fields, self._precision)

    def peer_swap(self, path, viewpoint, rank, world, n_fields=1):
        """The NVLink peer-memory compositor for this sampler's frame, in the sampler's partial format."""
        from .parallel import PeerSwap
        return PeerSwap(self._ctx, self._frame, path, viewpoint, rank, world, fmt=self.partial_format(n_fields))

    def flush_parallel(self, swap, broadcast=True):
        """Multi-GPU flush: cast this rank's bricks into an affine partial image, composite the partials
        of all ranks in view order (binary swap), leave the result in the host image of rank 0 (of every
        rank with ``broadcast``).  Replaces reduce_tree_images (amr_kdtree.py:297-323)."""
        import torch
        fr = self._frame
        fr.upload(self.aimage_host, self.azbuffer)
        work = list(self._batches)
        if self._pending:
            bs = build_brickset(self._pending, self._ctx, self._storage)
            work.append((bs, np.arange(len(self._pending), dtype=np.int64)))
        nx, ny = self.aimage_host.shape[:2]
        dev = torch.device("cuda", self._ctx.device)
        stream = torch.cuda.ExternalStream(self._ctx.stream, device=dev)
        if len(work) > 1:
            raise NotImplementedError("one brick set per parallel flush")
        if work:
            bs, order = work[0]
        else:
            bs, order = self._empty_set(), np.zeros(0, dtype=np.int64)
        from .parallel import PeerSwap
        if isinstance(swap, PeerSwap):
            # NVLink peer-memory route: no collective, the swap rounds read the partners' HBM directly
            swap.cast(fr, self.kind, bs, order, tf=self._tables, n_samples=self.n_samples, precision=self._precision)
            swap.composite(fr, broadcast=broadcast)
            swap.check()
        else:
            with torch.cuda.stream(stream):
                partial = torch.empty((nx * ny, 8), dtype=torch.float64, device=dev)
                fr.cast_partial(self.kind, bs, order, partial.data_ptr(), tf=self._tables, n_samples=self.n_samples,
                                precision=self._precision)
                swap.composite(partial, fr, broadcast=broadcast)
                stream.synchronize()
        if broadcast or swap.rank == 0:
            fr.download(self.aimage_host)
        self._pending, self._batches, self._dirty = [], [], False

    def _empty_set(self):
        """A committed one-brick set far outside any view box, for ranks that own no brick."""
        cen = L.CELL_CENTRED if self.kind == L.PROJECTION else L.VERTEX_CENTRED
        n = 1 if cen == L.CELL_CENTRED else 2
        nf = 1
        if self._tables is not None:
            nf = 1 + max(int(f) for f in list(self.tf_obj.field_ids)[: self.tf_obj.n_field_tables])
        pg = PartitionedGrid(-1, [np.zeros((n, n, n))] * nf, np.ones((1, 1, 1), "uint8"), np.full(3, 1e30), np.full(3, 2e30),
                             np.ones(3, "int64"))
        return build_brickset([pg], self._ctx, self._storage)

    @property
    def aimage(self):
        self.flush()
        return self.aimage_host

    def count_samples(self, brickset, order):
        """Exact number of cells / samples the cast visits (the Gsamples/s denominator)."""
        self._frame.upload(self.aimage_host, self.azbuffer)
        st = self._frame.count(self.kind, brickset, order, self.n_samples)
        return int(st.n_cells), int(st.n_samples)

    def ensure_code_unit_params(self, params):
        for param_name in ["center", "vp_pos", "vp_dir", "width"]:
            param = params[param_name]
            if hasattr(param, "in_units"):
                params[param_name] = param.in_units("code_length")
        bounds = params["bounds"]
        if hasattr(bounds[0], "units"):
            params["bounds"] = tuple(b.in_units("code_length").d for b in bounds)
        return params


class ProjectionSampler(ImageSampler):
    """image_samplers.pyx:325-343; the extra positional ``num_samples`` of utils.py:129 is swallowed."""
    kind = L.PROJECTION


class InterpolatedProjectionSampler(ImageSampler):
    """image_samplers.pyx:346-402"""
    kind = L.INTERPOLATED_PROJECTION

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, n_samples=10, **kwargs):
        super().__init__(vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, **kwargs)
        self.n_samples = int(n_samples)


class VolumeRenderSampler(ImageSampler):
    """image_samplers.pyx:405-494"""
    kind = L.VOLUME_RENDER

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, tf_obj,
                 n_samples=10, **kwargs):
        super().__init__(vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, **kwargs)
        self.n_samples = int(n_samples)
        self.tf_obj = tf_obj
        self.my_field_tables = [(tf_obj.tables[i], tf_obj.tables[i].y) for i in range(tf_obj.n_field_tables)]
        self._tables = D.TransferTables(self._ctx, tf_obj)


class LightSourceRenderSampler(VolumeRenderSampler):
    """image_samplers.pyx:496-600: the volume render lit by one directional light (shading by the normalised gradient of
    field 0).  Same constructor as the reference, including the in-place normalisation of ``light_dir``."""

    def __init__(self, vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, tf_obj,
                 n_samples=10, light_dir=(1., 1., 1.), light_rgba=(1., 1., 1., 1.), **kwargs):
        super().__init__(vp_pos, vp_dir, center, bounds, image, x_vec, y_vec, width, volume_method, tf_obj, n_samples, **kwargs)
        light_dir /= np.sqrt(light_dir[0] * light_dir[0] + light_dir[1] * light_dir[1] + light_dir[2] * light_dir[2])
        self.light_dir = np.ascontiguousarray(light_dir, dtype=np.float64)
        self.light_rgba = np.ascontiguousarray(light_rgba, dtype=np.float64)
        if self.light_dir.shape != (3,) or self.light_rgba.shape != (4,):
            raise ValueError("light_dir needs 3 components and light_rgba 4")
        L.check(L.load().ytb200_tf_set_light(self._tables.handle, self.light_dir.ctypes.data_as(L.dp),
                                             self.light_rgba.ctypes.data_as(L.dp)))
