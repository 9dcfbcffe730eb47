"""ctypes binding of ``libytb200.so`` (C ABI declared in ``include/ytb200.h``).

The product path has no CPU fallback: if the shared library is missing or no CUDA device is
visible, the first call that needs it raises.  Building is done by ``yt_b200.build.build()``
(nvcc, sm_100a) -- never implicitly at import time.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YTB200_LIB") or os.path.join(HERE, "libytb200.so")  # override: kernel experiments only

PROJECTION, INTERPOLATED_PROJECTION, VOLUME_RENDER = 0, 1, 2
LENS_PLANE_PARALLEL, LENS_PERSPECTIVE, LENS_GENERIC = 0, 1, 2
CELL_CENTRED, VERTEX_CENTRED = 0, 1
STORE_F32, STORE_F64 = 0, 1
PREC_FAST, PREC_EXACT = 0, 1
# partial-image formats (include/ytb200.h: YTB200_PARTIAL_*)
PARTIAL_AB64, PARTIAL_AB32, PARTIAL_A32, PARTIAL_A32G, PARTIAL_SUM1, PARTIAL_SUM2, PARTIAL_SUM4 = range(7)

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


class Camera(C.Structure):
    """struct ytb200_camera"""
    _fields_ = [("lens", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
                ("vp_pos", C.c_double * 12), ("vp_dir", C.c_double * 3),
                ("vp_pos_px", dp), ("vp_dir_px", dp),
                ("center", C.c_double * 3), ("bounds", C.c_double * 4),
                ("x_vec", C.c_double * 3), ("y_vec", C.c_double * 3), ("width", C.c_double * 3),
                ("has_camera_data", C.c_int32), ("camera_data", C.c_double * 15)]


class Stats(C.Structure):
    """struct ytb200_stats"""
    _fields_ = [("kernel_ms", C.c_double), ("n_launches", C.c_int64),
                ("n_cells", C.c_int64), ("n_samples", C.c_int64)]


class YTB200Error(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/ytb200.h declares
SIGNATURES = {
    "ytb200_abi_version": (C.c_int, []),
    "ytb200_last_error": (C.c_char_p, []),
    "ytb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ytb200_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ytb200_ctx_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "ytb200_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "ytb200_ctx_sm_count": (C.c_int, [C.c_void_p]),
    "ytb200_bricks_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ytb200_brick_upload": (C.c_int, [C.c_void_p, C.c_int64, i64p, dp, dp, C.POINTER(dp), u8p]),
    "ytb200_brick_attach": (C.c_int, [C.c_void_p, C.c_int64, i64p, dp, dp, C.POINTER(C.c_void_p), C.c_void_p]),
    "ytb200_cells_upload": (C.c_int, [C.c_void_p, dp, i64p, C.POINTER(C.c_void_p)]),
    "ytb200_cells_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ytb200_brick_from_cells": (C.c_int, [C.c_void_p, C.c_int64, i64p, i64p, i64p, dp, dp, C.POINTER(C.c_void_p), C.c_int,
                                          i32p, u8p, C.POINTER(Stats)]),
    "ytb200_brick_from_ghost_cells": (C.c_int, [C.c_void_p, C.c_int64, i64p, i64p, i64p, dp, dp, C.POINTER(C.c_void_p), C.c_int,
                                                i32p, u8p, C.POINTER(Stats)]),
    "ytb200_brick_attach_pitched": (C.c_int, [C.c_void_p, C.c_int64, i64p, dp, dp, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64]),
    "ytb200_brick_pitch": (C.c_int, [C.c_void_p, C.c_int64, i64p, i64p]),
    "ytb200_bricks_commit": (C.c_int, [C.c_void_p]),
    "ytb200_bricks_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_brick_device_ptrs": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "ytb200_bricks_device_bytes": (C.c_int64, [C.c_void_p]),
    "ytb200_tf_create": (C.c_int, [C.c_void_p, C.c_int, i32p, C.POINTER(dp), dp, i32p, i32p, i32p, i32p, C.c_int,
                                   C.POINTER(C.c_void_p)]),
    "ytb200_tf_set_light": (C.c_int, [C.c_void_p, dp, dp]),
    "ytb200_tf_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_frame_create": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(C.c_void_p)]),
    "ytb200_frame_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_frame_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ytb200_frame_clear": (C.c_int, [C.c_void_p]),
    "ytb200_frame_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, i64p, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(Stats)]),
    "ytb200_frame_count": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, i64p, C.c_int64, C.c_int, C.POINTER(Stats)]),
    "ytb200_frame_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ytb200_frame_device_image": (C.c_void_p, [C.c_void_p]),
    "ytb200_frame_cast_partial": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, i64p, C.c_int64, C.c_void_p, C.c_int,
                                            C.c_int, C.c_void_p, C.POINTER(Stats)]),
    "ytb200_frame_cast_partial_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, i64p, C.c_int64, C.c_void_p, C.c_int,
                                                C.c_int, C.c_void_p, C.POINTER(Stats)]),
    "ytb200_frame_bind_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ytb200_frame_set_early_termination": (C.c_int, [C.c_void_p, C.c_double]),
    "ytb200_tf_line_form_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "ytb200_swap_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ytb200_swap_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_swap_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ytb200_swap_open": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ytb200_swap_plan": (C.c_int, [C.c_void_p, C.c_int, i32p, i64p, i64p, i64p, i64p, i32p]),
    "ytb200_swap_format": (C.c_int, [C.c_void_p]),
    "ytb200_partial_format": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "ytb200_partial_bytes": (C.c_int, [C.c_int]),
    "ytb200_swap_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, i64p, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_int, C.c_void_p]),
    "ytb200_swap_composite": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ytb200_swap_composite_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ytb200_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "ytb200_host_unregister": (C.c_int, [C.c_void_p]),
    "ytb200_swap_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ytb200_partial_compose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "ytb200_partial_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "ytb200_render": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, i64p, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "ytb200_kdtree_last_error": (C.c_char_p, []),
    "ytb200_kdtree_create": (C.c_int, [dp, dp, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ytb200_kdtree_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_kdtree_add_grids": (C.c_int, [C.c_void_p, C.c_int64, dp, dp, i64p]),
    "ytb200_kdtree_node_count": (C.c_int64, [C.c_void_p]),
    "ytb200_kdtree_nodes": (C.c_int, [C.c_void_p, i64p, i64p, i32p, i32p, i32p, dp, dp, dp]),
    "ytb200_kdtree_traverse": (C.c_int64, [C.c_void_p, dp, i32p, C.c_int64]),
    "ytb200_zlines": (C.c_int, [dp, dp, C.c_int64, C.c_int64, i64p, i64p, dp, C.c_int64, dp, C.c_int64, C.c_int, C.c_int,
                                C.c_int, C.c_int]),
    "ytb200_zpoints": (C.c_int, [dp, dp, C.c_int64, C.c_int64, i64p, i64p, dp, C.c_int64, i64p, dp, C.c_int64, i64p,
                                 C.c_int, C.c_int]),
    "ytb200_octree_create": (C.c_int, [C.c_void_p, C.c_int, dp, dp, C.POINTER(C.c_void_p)]),
    "ytb200_octree_destroy": (C.c_int, [C.c_void_p]),
    "ytb200_octree_add_nodes": (C.c_int, [C.c_void_p, C.c_int64, i32p, i32p, i32p]),
    "ytb200_octree_node_count": (C.c_int64, [C.c_void_p]),
    "ytb200_octree_set_cells": (C.c_int, [C.c_void_p, C.c_int64, dp]),
    "ytb200_frame_cast_octree": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Stats)]),
    "ytb200_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "ytb200_host_free": (C.c_int, [C.c_void_p]),
}

_lib = None


def load():
    """dlopen libytb200.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YTB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  yt_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("YTB200_LIB") and not hasattr(lib, name):
                continue   # kernel-experiment builds of an older source may lack newer entry points
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().ytb200_last_error()
        raise YTB200Error(f"libytb200 error {rc}: {msg.decode() if msg else '?'}")


def device_count():
    n = C.c_int(0)
    check(load().ytb200_device_count(C.byref(n)))
    return n.value
