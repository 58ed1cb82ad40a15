"""ctypes front end of oracle/_ref/libref_shader.so: the reference's SHIPPED compute shader
(assets/shaders/compute_pass.comp.spv) translated to C++ by oracle/spirv_to_cpp.py and compiled
for the host — the reference's own implementation of the hot path running on the CPU. TEST
INFRASTRUCTURE: a second checker next to the interpreter (it is ~100x faster, so full-size frames
can be compared live, also on the GPU box, where the prebuilt library travels) and the CPU
baseline of kind "reference" in bench.py. Built only next to the reference tree
(`make -C oracle ref_shader`); nothing here is committed source of the reference."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_ref" / "libref_shader.so"
REFERENCE_SPV = Path("/root/reference/assets/shaders/compute_pass.comp.spv")

_lib = None


def build() -> Path:
    """Translate + compile (needs the reference tree)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "ref_shader"], check=True)
    return LIB_PATH


def available() -> bool:
    return LIB_PATH.exists() or REFERENCE_SPV.exists()


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if REFERENCE_SPV.exists():
            build()  # make: no-op when up to date
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing and the reference tree is not here to build it")
        lib = C.CDLL(str(LIB_PATH))
        lib.rvpt_ref_shader_create.restype = C.c_void_p
        lib.rvpt_ref_shader_destroy.argtypes = [C.c_void_p]
        lib.rvpt_ref_shader_bind_buffer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        lib.rvpt_ref_shader_bind_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.rvpt_ref_shader_dispatch.restype = C.c_int
        lib.rvpt_ref_shader_dispatch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        lib.rvpt_ref_shader_hardware_threads.restype = C.c_int
        _lib = lib
    return _lib


class RefShaderRenderer:
    """Same constructor / render_frame signature as oracle.OracleRenderer and spirv_vm.SpirvRenderer."""

    def __init__(self, width: int, height: int, triangles, materials, nodes, unorm8: bool = False,
                 reference_dispatch: bool = False, nthreads: int = 0):
        self.lib = load()
        self.h = self.lib.rvpt_ref_shader_create()
        if not self.h:
            raise RuntimeError("ref_shader: module does not fit the runtime context")
        self.W, self.H, self.unorm8, self.nthreads = int(width), int(height), unorm8, nthreads
        self.ref_dispatch = reference_dispatch
        self.tris = np.ascontiguousarray(triangles)
        self.mats = np.ascontiguousarray(materials)
        self.nodes = np.ascontiguousarray(nodes)
        dt = np.uint8 if unorm8 else np.float32
        self.temporal = np.zeros((self.H, self.W, 4), dt)
        self.result = np.zeros((self.H, self.W, 4), dt)
        self.random = np.zeros(4, np.float32)
        b = self.lib.rvpt_ref_shader_bind_buffer
        b(self.h, 5, self.nodes.ctypes.data, self.nodes.nbytes)
        b(self.h, 6, self.tris.ctypes.data, self.tris.nbytes)
        b(self.h, 7, self.mats.ctypes.data, self.mats.nbytes)
        b(self.h, 3, self.random.ctypes.data, self.random.nbytes)
        self.lib.rvpt_ref_shader_bind_image(self.h, 1, self.result.ctypes.data, self.W, self.H, int(unorm8))
        self.lib.rvpt_ref_shader_bind_image(self.h, 2, self.temporal.ctypes.data, self.W, self.H, int(unorm8))

    def render_frame(self, settings, camera, y_begin: int = 0, y_end: int | None = None) -> None:
        self._rs = np.ascontiguousarray(settings)
        self._cam = np.ascontiguousarray(camera, np.float32)
        assert self._rs.dtype.itemsize == 40 and self._cam.size == 20
        self.lib.rvpt_ref_shader_bind_buffer(self.h, 0, self._rs.ctypes.data, 40)
        self.lib.rvpt_ref_shader_bind_buffer(self.h, 4, self._cam.ctypes.data, 80)
        W_eff, H_eff = self.W, self.H
        if self.ref_dispatch:
            W_eff, H_eff = (self.W // 16) * 16, (self.H // 16) * 16
        y1 = H_eff if y_end is None else min(y_end, H_eff)
        self.lib.rvpt_ref_shader_dispatch(self.h, 0, W_eff, y_begin, y1, self.nthreads)

    def result_rgba8(self) -> np.ndarray:
        """The result image as rgba8 codes (in float mode: rv_unorm8_store of the float image —
        NaN -> 0, clamp, round-to-nearest-even of x * 255 in float32)."""
        if self.unorm8:
            return self.result
        x = self.result
        c = np.where(x > 0, x, np.float32(0)).astype(np.float32)
        c = np.minimum(c, np.float32(1))
        return np.rint(c * np.float32(255)).astype(np.uint8)

    def threads(self) -> int:
        return self.nthreads or int(self.lib.rvpt_ref_shader_hardware_threads())

    def close(self):
        if self.h:
            self.lib.rvpt_ref_shader_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
