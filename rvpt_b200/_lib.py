"""ctypes binding of include/rvpt_abi.h (the C-ABI drop-in boundary).

The native library is built in-tree (`python -m rvpt_b200.build`) and MUST be
present: there is no Python / CPU fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# RVPT_B200_LIB: developer knob to A/B another build of the same library
LIB_PATH = Path(os.environ.get("RVPT_B200_LIB") or Path(__file__).resolve().parent / "librvpt_b200.so")

MAX_BOUNCE_STATS = 64
FLAG_ACCUM_RGBA8 = 0x1
FLAG_REFERENCE_DISPATCH = 0x2
FLAG_BRUTE_FORCE = 0x4
FLAG_UNFUSED = 0x8
FLAG_NO_OCTANTS = 0x10
FLAG_NO_FORECAST = 0x40
FLAG_REFERENCE_ORDER = 0x80
FLAG_NO_QUEUE_SORT = 0x100
FLAG_NO_BATCH = 0x200
FLAG_GPU_BVH = 0x400
FLAG_NO_LEAF_LISTS = 0x800

OK, EINVAL, ECUDA, ENOSCENE, EUNSUPPORTED, ENOMEM = 0, -1, -2, -3, -4, -5


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64),
        ("segments", C.c_uint64),
        ("active", C.c_uint64 * MAX_BOUNCE_STATS),
        ("kernel_launches", C.c_uint32),
        ("traversal_order", C.c_uint32),
        ("frames", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class KernelTimes(C.Structure):
    _fields_ = [
        ("primary_ms", C.c_double), ("bounce_ms", C.c_double),
        ("primary_launches", C.c_uint32), ("bounce_launches", C.c_uint32),
    ]


class TileInfo(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("tiles_x", C.c_uint32), ("tiles_y", C.c_uint32),
        ("rank", C.c_uint32), ("nranks", C.c_uint32),
        ("n_local_tiles", C.c_uint32), ("n_local_tiles_padded", C.c_uint32),
        ("d_accum_tiles", C.c_void_p), ("d_rgba8_tiles", C.c_void_p),
        ("accum_bytes", C.c_uint64), ("rgba8_bytes", C.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/rvpt_abi.h declares
SIGNATURES = {
    "rvpt_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rvpt_b200_destroy": (None, [C.c_void_p]),
    "rvpt_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "rvpt_b200_set_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rvpt_b200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_upload_scene": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_size_t]),
    "rvpt_b200_render_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rvpt_b200_render_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "rvpt_b200_sync": (C.c_int, [C.c_void_p]),
    "rvpt_b200_read_output_rgba8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_read_accum_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_write_accum_f32": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_reset_accum": (C.c_int, [C.c_void_p]),
    "rvpt_b200_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "rvpt_b200_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "rvpt_b200_get_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(KernelTimes)]),
    "rvpt_b200_set_timeline": (C.c_int, [C.c_void_p, C.c_int]),
    "rvpt_b200_get_timeline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32)]),
    "rvpt_b200_get_tile_info": (C.c_int, [C.c_void_p, C.POINTER(TileInfo)]),
    "rvpt_b200_set_external_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rvpt_b200_untile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "rvpt_b200_untile_on": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                      C.c_void_p]),
    "rvpt_b200_export_output": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_attach_output": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_read_output_rgba8_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_wait_output": (C.c_int, [C.c_void_p]),
    "rvpt_b200_flip_output": (C.c_int, [C.c_void_p]),
    "rvpt_b200_export_output2": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_attach_output2": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rvpt_b200_build_bvh": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t),
                                      C.c_void_p]),
    "rvpt_b200_build_bvh_gpu": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t),
                                          C.c_void_p, C.POINTER(C.c_float)]),
    "rvpt_b200_camera_data": (None, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p]),
    "rvpt_b200_has_coincident_faces": (C.c_int, [C.c_void_p, C.c_size_t]),
    "rvpt_b200_octant_layouts": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                           C.c_size_t, C.POINTER(C.c_size_t)]),
    "rvpt_b200_abi_version": (C.c_uint32, []),
    "rvpt_b200_build_info": (C.c_char_p, []),
    "rvpt_b200_selftest_math": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the native library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m rvpt_b200.build` "
                "(there is no CPU fallback for the render path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
