"""ctypes wrapper of the CPU oracle (oracle/rvpt_oracle.cpp).

TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and the CPU-baseline
legs of bench.py import this; the product package (rvpt_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "librvpt_oracle.so"

FLAG_ACCUM_RGBA8 = 0x1
FLAG_REFERENCE_DISPATCH = 0x2
FLAG_BRUTE_FORCE = 0x4

_lib = None

# RVPT::RenderSettings, src/rvpt/rvpt.h:77-89 (the oracle's own copy: bench.py's reference arm
# must not import the product package)
RENDER_SETTINGS_DTYPE = np.dtype([
    ("max_bounces", "<i4"), ("aa", "<i4"), ("current_frame", "<u4"), ("camera_mode", "<i4"),
    ("top_left_render_mode", "<i4"), ("top_right_render_mode", "<i4"),
    ("bottom_left_render_mode", "<i4"), ("bottom_right_render_mode", "<i4"), ("split_ratio", "<f4", 2)])
assert RENDER_SETTINGS_DTYPE.itemsize == 40


def settings(max_bounces: int = 8, aa: int = 1, frame: int = 0, camera_mode: int = 0, mode: int = 9) -> np.ndarray:
    rs = np.zeros(1, RENDER_SETTINGS_DTYPE)
    rs["max_bounces"], rs["aa"], rs["current_frame"], rs["camera_mode"] = max_bounces, aa, frame, camera_mode
    for k in ("top_left", "top_right", "bottom_left", "bottom_right"):
        rs[f"{k}_render_mode"] = mode
    rs["split_ratio"] = (0.5, 0.5)
    return rs


def load_workload(name: str) -> dict:
    """A committed bench workload (oracle/workloads/, written by tools/make_bench_workloads.py):
    nodes, triangles (BVH order), materials, camera_16x9."""
    with np.load(HERE / "workloads" / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def build(force: bool = False) -> Path:
    src = HERE / "rvpt_oracle.cpp"
    deps = [src, HERE.parent / "include" / "rvpt_abi.h", HERE.parent / "include" / "rvpt_math.h"]
    stale = (not LIB_PATH.exists()) or any(d.stat().st_mtime > LIB_PATH.stat().st_mtime for d in deps)
    if force or stale:
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB_PATH))
        lib.rvpt_oracle_render_rows.restype = C.c_int
        lib.rvpt_oracle_render_rows.argtypes = [
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
            C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.rvpt_oracle_wang_hash.restype = C.c_uint32
        lib.rvpt_oracle_wang_hash.argtypes = [C.c_uint32]
        lib.rvpt_oracle_rand_stream.restype = None
        lib.rvpt_oracle_rand_stream.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.c_void_p, C.c_void_p]
        lib.rvpt_oracle_sincos.restype = None
        lib.rvpt_oracle_sincos.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.rvpt_oracle_normalize.restype = None
        lib.rvpt_oracle_normalize.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.rvpt_oracle_camera_ray.restype = None
        lib.rvpt_oracle_camera_ray.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        lib.rvpt_oracle_intersect_triangle.restype = C.c_int
        lib.rvpt_oracle_intersect_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                       C.c_void_p]
        lib.rvpt_oracle_intersect_aabb.restype = C.c_int
        lib.rvpt_oracle_intersect_aabb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                                   C.c_float]
        lib.rvpt_oracle_intersect_scene.restype = C.c_int64
        lib.rvpt_oracle_intersect_scene.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                    C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                                    C.c_size_t, C.c_void_p]
        lib.rvpt_oracle_fresnel.restype = C.c_float
        lib.rvpt_oracle_fresnel.argtypes = [C.c_float, C.c_float, C.c_float]
        lib.rvpt_oracle_contract_probe.restype = C.c_int
        lib.rvpt_oracle_hardware_threads.restype = C.c_int
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class OracleRenderer:
    """Progressive renderer with the reference's per-frame semantics
    (compute_pass.comp:121-167). State = the temporal image (float32 running
    mean, or the rgba8 image in ACCUM_RGBA8 mode) + the result image."""

    def __init__(self, width: int, height: int, triangles, materials, nodes=None, flags: int = 0,
                 nthreads: int = 0):
        self.lib = load()
        self.W, self.H, self.flags, self.nthreads = int(width), int(height), int(flags), nthreads
        self.tris = np.ascontiguousarray(triangles)
        self.mats = np.ascontiguousarray(materials)
        self.nodes = None if nodes is None else np.ascontiguousarray(nodes)
        assert self.tris.dtype.itemsize == 64 and self.mats.dtype.itemsize == 48
        self.accum = np.zeros((self.H, self.W, 4), np.float32)
        self.temporal = np.zeros((self.H, self.W, 4), np.uint8)
        self.result = np.zeros((self.H, self.W, 4), np.uint8)
        self.active = np.zeros(64, np.uint64)

    def render_frame(self, settings, camera, y_begin: int = 0, y_end: int | None = None) -> None:
        rs = np.ascontiguousarray(settings)
        cam = np.ascontiguousarray(camera, np.float32)
        assert rs.dtype.itemsize == 40 and cam.size == 20
        self.active[:] = 0
        rc = self.lib.rvpt_oracle_render_rows(
            _ptr(self.nodes), 0 if self.nodes is None else len(self.nodes),
            _ptr(self.tris), len(self.tris), _ptr(self.mats), len(self.mats),
            _ptr(rs), _ptr(cam), self.W, self.H, self.flags,
            y_begin, self.H if y_end is None else y_end,
            _ptr(self.accum), _ptr(self.temporal), _ptr(self.result), _ptr(self.active),
            self.nthreads)
        if rc:
            raise RuntimeError(f"oracle error {rc}")

    def accum_f32(self) -> np.ndarray:
        if self.flags & FLAG_ACCUM_RGBA8:
            out = self.temporal.astype(np.float32) / np.float32(255.0)
            out[..., 3] = 0
            return out
        return self.accum

    def active_list(self) -> list[int]:
        a = [int(v) for v in self.active]
        while a and a[-1] == 0:
            a.pop()
        return a
