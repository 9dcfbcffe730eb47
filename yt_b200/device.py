"""Thin Python handles over the C-ABI objects of libytb200 (ctx, brick set, tables, frame).

Nothing here computes: these classes marshal numpy arrays into the plain-pointer calls of
``include/ytb200.h`` and own the opaque handles (every create has a destroy).
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib as L

_contexts = {}


class Context:
    """One device + one stream (``ytb200_ctx``)."""

    def __init__(self, device=0, stream=None):
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.ytb200_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = int(device)
        self._fin = weakref.finalize(self, lib.ytb200_ctx_destroy, h)

    def synchronize(self):
        L.check(L.load().ytb200_ctx_synchronize(self.handle))

    @property
    def stream(self):
        return L.load().ytb200_ctx_stream(self.handle)

    @property
    def sm_count(self):
        return L.load().ytb200_ctx_sm_count(self.handle)


def default_device():
    """Device index for this process: ``YTB200_DEVICE`` or ``LOCAL_RANK`` (torchrun) or 0."""
    for k in ("YTB200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(k, "") != "":
            return int(os.environ[k])
    return 0


def get_context(device=None):
    """Process-wide default context per device (created on first use)."""
    if device is None:
        device = default_device()
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dptr(a):
    return a.ctypes.data_as(L.dp)


class DeviceCells:
    """One grid's cell-centred fp64 field resident on the device (``ytb200_cells_upload``): the input of
    the on-device vertex centring.  Uploaded once per (grid, field); every kd leaf of the grid reads it."""

    def __init__(self, ctx, cells):
        a = _f64(cells)
        if a.ndim != 3:
            raise ValueError("cells must be a 3-D array")
        dims = np.array(a.shape, dtype=np.int64)
        p = C.c_void_p()
        lib = L.load()
        L.check(lib.ytb200_cells_upload(ctx.handle, _dptr(a), dims.ctypes.data_as(L.i64p), C.byref(p)))
        self.ptr, self.shape, self.ctx = int(p.value), tuple(int(d) for d in a.shape), ctx
        self._fin = weakref.finalize(self, lib.ytb200_cells_free, ctx.handle, C.c_void_p(self.ptr))


class BrickSet:
    """Device-resident list of bricks (``ytb200_brickset``)."""

    def __init__(self, ctx, n_bricks, n_fields, centring, storage=L.STORE_F32):
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.ytb200_bricks_create(ctx.handle, int(n_bricks), int(n_fields), int(centring), int(storage), C.byref(h)))
        self.handle, self.ctx = h, ctx
        self.n, self.n_fields, self.centring, self.storage = int(n_bricks), int(n_fields), int(centring), int(storage)
        self._keep = []      # objects whose device memory we borrow (attach)
        self._fin = weakref.finalize(self, lib.ytb200_bricks_destroy, h)

    def upload(self, index, data, mask, left_edge, right_edge, dims):
        """Host fp64 arrays -> device (converted to the set's storage type)."""
        dims = np.ascontiguousarray(dims, dtype=np.int64)
        c = 1 if self.centring == L.VERTEX_CENTRED else 0
        shape = tuple(int(d) + c for d in dims)
        arrs = []
        for f in range(self.n_fields):
            a = _f64(data[f])
            if a.shape != shape:
                raise ValueError(f"field {f} has shape {a.shape}, expected {shape}")
            arrs.append(a)
        ptrs = (L.dp * self.n_fields)(*[_dptr(a) for a in arrs])
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            if m.shape != tuple(int(d) for d in dims):
                raise ValueError(f"mask has shape {m.shape}, expected {tuple(int(d) for d in dims)}")
        le, re = _f64(left_edge), _f64(right_edge)
        L.check(L.load().ytb200_brick_upload(
            self.handle, int(index), dims.ctypes.data_as(L.i64p), _dptr(le), _dptr(re), ptrs,
            m.ctypes.data_as(L.u8p) if m is not None else None))

    def from_cells(self, index, cells_ptrs, grid_dims, lo, hi, left_edge, right_edge, log_fields=None, mask=None,
                   cells_are_f32=False, keep=None, timed=False, ghost=False):
        """Brick `index` := vertices lo..hi (inclusive) of the grid whose cell-centred fields live at the device
        pointers `cells_ptrs`: no-ghost vertex centring + log10 + slice + storage conversion in one kernel
        (grid_patch.py:284-296, amr_kdtree.py:343-365).  ``ghost``: `cells_ptrs` hold the grid padded with one layer of
        ghost zones (shape grid_dims + 2) and vertices are 8-cell means (grid_patch.py:298-310)."""
        gd = np.ascontiguousarray(grid_dims, dtype=np.int64)
        lo = np.ascontiguousarray(lo, dtype=np.int64)
        hi = np.ascontiguousarray(hi, dtype=np.int64)
        ptrs = (C.c_void_p * self.n_fields)(*[C.c_void_p(int(p)) for p in cells_ptrs])
        lg = None
        if log_fields is not None:
            lg = np.ascontiguousarray([1 if x else 0 for x in log_fields], dtype=np.int32)
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            if m.shape != tuple(int(h - l) for l, h in zip(lo, hi)):
                raise ValueError(f"mask has shape {m.shape}, expected {tuple(int(h - l) for l, h in zip(lo, hi))}")
        le, re = _f64(left_edge), _f64(right_edge)
        st = L.Stats() if timed else None
        fn = L.load().ytb200_brick_from_ghost_cells if ghost else L.load().ytb200_brick_from_cells
        L.check(fn(
            self.handle, int(index), gd.ctypes.data_as(L.i64p), lo.ctypes.data_as(L.i64p), hi.ctypes.data_as(L.i64p),
            _dptr(le), _dptr(re), ptrs, 1 if cells_are_f32 else 0, lg.ctypes.data_as(L.i32p) if lg is not None else None,
            m.ctypes.data_as(L.u8p) if m is not None else None, C.byref(st) if timed else None))
        if keep is not None:
            self._keep.append(keep)
        return st

    def attach(self, index, dev_ptrs, dev_mask, left_edge, right_edge, dims, keep=None, row_pitch=None):
        """Borrow device arrays (already in this set's storage type); `keep` is kept alive.  ``row_pitch``: elements
        between consecutive z-rows when the arrays are padded (default: contiguous)."""
        dims = np.ascontiguousarray(dims, dtype=np.int64)
        ptrs = (C.c_void_p * self.n_fields)(*[C.c_void_p(int(p)) for p in dev_ptrs])
        le, re = _f64(left_edge), _f64(right_edge)
        if row_pitch is None:
            L.check(L.load().ytb200_brick_attach(
                self.handle, int(index), dims.ctypes.data_as(L.i64p), _dptr(le), _dptr(re), ptrs,
                C.c_void_p(int(dev_mask)) if dev_mask else None))
        else:
            L.check(L.load().ytb200_brick_attach_pitched(
                self.handle, int(index), dims.ctypes.data_as(L.i64p), _dptr(le), _dptr(re), ptrs,
                C.c_void_p(int(dev_mask)) if dev_mask else None, int(row_pitch)))
        if keep is not None:
            self._keep.append(keep)

    def commit(self):
        L.check(L.load().ytb200_bricks_commit(self.handle))
        return self

    def device_ptrs(self, index):
        data = (C.c_void_p * self.n_fields)()
        mask = C.c_void_p()
        L.check(L.load().ytb200_brick_device_ptrs(self.handle, int(index), data, C.byref(mask)))
        return [int(p or 0) for p in data], int(mask.value or 0)

    def pitch(self, index):
        """(plane pitch, row pitch) of brick `index`'s arrays, in elements."""
        a, b = C.c_int64(0), C.c_int64(0)
        L.check(L.load().ytb200_brick_pitch(self.handle, int(index), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    @property
    def device_bytes(self):
        return L.load().ytb200_bricks_device_bytes(self.handle)


class TransferTables:
    """Device-resident FieldInterpolationTables built from a transfer-function object with the
    reference's duck type (image_samplers.pyx:429-446): ``n_field_tables``, ``tables[i].y``,
    ``tables[i].x_bounds``, ``field_ids``, ``weight_field_ids``, ``weight_table_ids``,
    ``field_table_ids``, ``grey_opacity``."""

    def __init__(self, ctx, tf_obj):
        lib = L.load()
        n = int(tf_obj.n_field_tables)
        assert n <= 6
        ys = [_f64(tf_obj.tables[i].y) for i in range(n)]
        nbins = np.array([y.shape[0] for y in ys], dtype=np.int32)
        bounds = np.array([[float(tf_obj.tables[i].x_bounds[0]), float(tf_obj.tables[i].x_bounds[1])]
                           for i in range(n)], dtype=np.float64)
        yp = (L.dp * n)(*[_dptr(y) for y in ys])

        def i32(v):
            return np.array([int(x) for x in list(v)[:6]], dtype=np.int32)
        fid, wfid, wtid, ftid = i32(tf_obj.field_ids), i32(tf_obj.weight_field_ids), i32(tf_obj.weight_table_ids), i32(tf_obj.field_table_ids)
        h = C.c_void_p()
        L.check(lib.ytb200_tf_create(
            ctx.handle, n, nbins.ctypes.data_as(L.i32p), yp, _dptr(bounds), fid.ctypes.data_as(L.i32p),
            wfid.ctypes.data_as(L.i32p), wtid.ctypes.data_as(L.i32p), ftid.ctypes.data_as(L.i32p),
            int(bool(getattr(tf_obj, "grey_opacity", 0))), C.byref(h)))
        self.handle, self.ctx = h, ctx
        self.grey_opacity = bool(getattr(tf_obj, "grey_opacity", 0))
        self._fin = weakref.finalize(self, lib.ytb200_tf_destroy, h)


class Frame:
    """Device image + z-buffer + ray set of one sampler (``ytb200_frame``)."""

    def __init__(self, ctx, cam_struct, keep=()):
        lib = L.load()
        h = C.c_void_p()
        L.check(lib.ytb200_frame_create(ctx.handle, C.byref(cam_struct), C.byref(h)))
        self.handle, self.ctx = h, ctx
        self.nx, self.ny = cam_struct.nx, cam_struct.ny
        self._keep = keep
        self._fin = weakref.finalize(self, lib.ytb200_frame_destroy, h)

    def upload(self, image, zbuffer=None):
        assert image.dtype == np.float64 and image.flags.c_contiguous
        L.check(L.load().ytb200_frame_upload(
            self.handle, image.ctypes.data, zbuffer.ctypes.data if zbuffer is not None else None))

    def clear(self):
        L.check(L.load().ytb200_frame_clear(self.handle))

    def bind_host(self, image=None, zbuffer=None):
        """Multi-GPU frames: read the initial colours / z-buffer from these PAGE-LOCKED host arrays in place
        (``ytb200_frame_bind_host``); ``image=None`` unbinds."""
        L.check(L.load().ytb200_frame_bind_host(
            self.handle, image.ctypes.data if image is not None else None,
            zbuffer.ctypes.data if (image is not None and zbuffer is not None) else None))
        self._bound = (image, zbuffer)

    def set_early_termination(self, eps):
        """Early ray termination for grey-opacity ColorTransferFunction renders (``ytb200_frame_set_early_termination``);
        0 disables."""
        L.check(L.load().ytb200_frame_set_early_termination(self.handle, float(eps)))

    def _order(self, order):
        o = np.ascontiguousarray(order, dtype=np.int64)
        return o, o.ctypes.data_as(L.i64p), int(o.shape[0])

    def cast(self, kind, bricks, order, tf=None, n_samples=10, precision=L.PREC_FAST, timed=False):
        o, op, n = self._order(order)
        st = L.Stats() if timed else None
        L.check(L.load().ytb200_frame_cast(
            self.handle, int(kind), bricks.handle, op, n, tf.handle if tf is not None else None,
            int(n_samples), int(precision), C.byref(st) if timed else None))
        return st

    def cast_partial(self, kind, bricks, order, partial_ptr, tf=None, n_samples=10, precision=L.PREC_FAST, timed=False,
                     f32=False):
        o, op, n = self._order(order)
        st = L.Stats() if timed else None
        fn = L.load().ytb200_frame_cast_partial_f32 if f32 else L.load().ytb200_frame_cast_partial
        L.check(fn(
            self.handle, int(kind), bricks.handle, op, n, tf.handle if tf is not None else None,
            int(n_samples), int(precision), C.c_void_p(int(partial_ptr)), C.byref(st) if timed else None))
        return st

    def count(self, kind, bricks, order, n_samples=10):
        o, op, n = self._order(order)
        st = L.Stats()
        L.check(L.load().ytb200_frame_count(self.handle, int(kind), bricks.handle, op, n, int(n_samples), C.byref(st)))
        return st

    def download(self, image):
        assert image.dtype == np.float64 and image.flags.c_contiguous
        L.check(L.load().ytb200_frame_download(self.handle, image.ctypes.data))
        return image

    def render(self, kind, bricks, order, image, zbuffer=None, tf=None, n_samples=10, precision=L.PREC_FAST):
        """upload -> cast -> download through the single C call the e2e benchmark times."""
        o, op, n = self._order(order)
        L.check(L.load().ytb200_render(
            self.handle, int(kind), bricks.handle, op, n, tf.handle if tf is not None else None, int(n_samples),
            int(precision), zbuffer.ctypes.data if zbuffer is not None else None, image.ctypes.data, None))
        return image

    @property
    def device_image(self):
        return L.load().ytb200_frame_device_image(self.handle)


_pinned_pool = {}       # nbytes -> [address, ...] of page-locked buffers whose arrays have been collected
_PINNED_POOL_MAX = 4    # buffers kept per size


def _pinned_release(nbytes, addr):
    free = _pinned_pool.setdefault(nbytes, [])
    if len(free) < _PINNED_POOL_MAX:
        free.append(addr)
    else:
        L.load().ytb200_host_free(C.c_void_p(addr))


def pinned_empty(shape, dtype=np.float64):
    """numpy array over cudaHostAlloc'd memory.  When the array (and every view of it) is collected the buffer goes to
    a small per-size pool instead of back to the driver: page-locking 134 MB costs tens of milliseconds, and a render
    loop asks for the same image and z-buffer sizes every frame."""
    lib = L.load()
    dt = np.dtype(dtype)
    n = max(int(np.prod(shape)) * dt.itemsize, 1)
    free = _pinned_pool.setdefault(n, [])
    if not free:
        # page-locking is slow (~1.5 ms per MB): take image-sized buffers three at a time, so that a render loop whose
        # caller still holds the previous frame or two never waits for the driver again
        for _ in range(3 if n >= (8 << 20) else 1):
            p = C.c_void_p()
            L.check(lib.ytb200_host_alloc(n, C.byref(p)))
            free.append(p.value)
    addr = free.pop()
    buf = (C.c_char * n).from_address(addr)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, _pinned_release, n, addr)
    return arr


_fill_pool = None


def parallel_assign(view, value, threads=8):
    """``view[...] = value`` in slabs along axis 0 on a few threads (numpy releases the GIL in these loops).  A 2048^2 x 4
    fp64 image is 134 MB: zeroing it takes 9 ms on one core of the GPU box and 2 ms on eight, its alpha channel 13 ms
    against 2.5 -- together with the z-buffer fills that was 30 ms of host time in every Scene.render()."""
    global _fill_pool
    n = view.shape[0] if view.ndim else 0
    if view.size * view.itemsize < (4 << 20) or n < 2 * threads or threads < 2:
        view[...] = value
        return view
    if _fill_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _fill_pool = ThreadPoolExecutor(max_workers=threads, thread_name_prefix="ytb200-fill")
    bounds = [n * i // threads for i in range(threads + 1)]
    list(_fill_pool.map(lambda k: view[bounds[k]:bounds[k + 1]].__setitem__(Ellipsis, value), range(threads)))
    return view


def pinned_full(shape, value, dtype=np.float64):
    return parallel_assign(pinned_empty(shape, dtype), value)
