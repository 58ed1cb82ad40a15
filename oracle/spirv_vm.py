"""ctypes front end of oracle/spirv_vm.cpp: runs the reference's SHIPPED compute shader
(assets/shaders/compute_pass.comp.spv) on the CPU, one invocation per pixel, with the
reference's descriptor bindings. TEST INFRASTRUCTURE: it pins oracle/rvpt_oracle.cpp to an
artefact the reference holds (tests/test_spirv_pin.py, tests/golden/make_spirv_golden.py).
The .spv is read from the reference tree at run time and never copied into this repo."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "librvpt_spirv_vm.so"
REFERENCE_SPV = Path("/root/reference/assets/shaders/compute_pass.comp.spv")

_lib = None


def build(force: bool = False) -> Path:
    args = ["make", "-s", "-C", str(HERE), "vm"]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        src = HERE / "spirv_vm.cpp"
        if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
            build()
        lib = C.CDLL(str(LIB_PATH))
        lib.rvpt_spirv_vm_create.restype = C.c_void_p
        lib.rvpt_spirv_vm_create.argtypes = [C.c_void_p, C.c_size_t]
        lib.rvpt_spirv_vm_destroy.argtypes = [C.c_void_p]
        lib.rvpt_spirv_vm_error.restype = C.c_char_p
        lib.rvpt_spirv_vm_error.argtypes = [C.c_void_p]
        lib.rvpt_spirv_vm_bind_buffer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        lib.rvpt_spirv_vm_bind_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.rvpt_spirv_vm_dispatch.restype = C.c_int
        lib.rvpt_spirv_vm_dispatch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        lib.rvpt_spirv_vm_executed.restype = C.c_uint64
        lib.rvpt_spirv_vm_executed.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def available() -> bool:
    return REFERENCE_SPV.exists()


class SpirvRenderer:
    """The reference's render loop around its own shader binary: same constructor and
    render_frame signature as oracle.OracleRenderer.

    unorm8=True   the two storage images are rgba8 like the reference's (rvpt.cpp:759-766,
                  803-811): imageLoad returns k/255, imageStore rounds clamp(x,0,1)*255;
    unorm8=False  float32 images, what the engine's default float accumulation mode keeps.
    reference_dispatch=True covers floor(W/16) x floor(H/16) workgroups (rvpt.cpp:1035-1036)."""

    def __init__(self, width: int, height: int, triangles, materials, nodes, unorm8: bool = False,
                 reference_dispatch: bool = False, spv: Path = REFERENCE_SPV, nthreads: int = 0):
        self.lib = load()
        words = np.fromfile(str(spv), dtype="<u4")
        self.vm = self.lib.rvpt_spirv_vm_create(words.ctypes.data, len(words))
        self._check()
        self.W, self.H, self.unorm8, self.nthreads = int(width), int(height), unorm8, nthreads
        self.ref_dispatch = reference_dispatch
        self.tris = np.ascontiguousarray(triangles)
        self.mats = np.ascontiguousarray(materials)
        self.nodes = np.ascontiguousarray(nodes)
        assert self.tris.dtype.itemsize == 64 and self.mats.dtype.itemsize == 48 and self.nodes.dtype.itemsize == 32
        dt = np.uint8 if unorm8 else np.float32
        self.temporal = np.zeros((self.H, self.W, 4), dt)
        self.result = np.zeros((self.H, self.W, 4), dt)
        self.random = np.zeros(4, np.float32)  # binding 3: declared, never read
        b = self.lib.rvpt_spirv_vm_bind_buffer
        b(self.vm, 5, self.nodes.ctypes.data, self.nodes.nbytes)
        b(self.vm, 6, self.tris.ctypes.data, self.tris.nbytes)
        b(self.vm, 7, self.mats.ctypes.data, self.mats.nbytes)
        b(self.vm, 3, self.random.ctypes.data, self.random.nbytes)
        self.lib.rvpt_spirv_vm_bind_image(self.vm, 1, self.result.ctypes.data, self.W, self.H, int(unorm8))
        self.lib.rvpt_spirv_vm_bind_image(self.vm, 2, self.temporal.ctypes.data, self.W, self.H, int(unorm8))

    def _check(self):
        msg = self.lib.rvpt_spirv_vm_error(self.vm)
        if msg:
            raise RuntimeError("spirv_vm: " + msg.decode())

    def render_frame(self, settings, camera, y_begin: int = 0, y_end: int | None = None,
                     x_begin: int = 0, x_end: int | None = None) -> None:
        self._rs = np.ascontiguousarray(settings)
        self._cam = np.ascontiguousarray(camera, np.float32)
        assert self._rs.dtype.itemsize == 40 and self._cam.size == 20
        self.lib.rvpt_spirv_vm_bind_buffer(self.vm, 0, self._rs.ctypes.data, 40)
        self.lib.rvpt_spirv_vm_bind_buffer(self.vm, 4, self._cam.ctypes.data, 80)
        W_eff, H_eff = self.W, self.H
        if self.ref_dispatch:
            W_eff, H_eff = (self.W // 16) * 16, (self.H // 16) * 16
        x1 = W_eff if x_end is None else min(x_end, W_eff)
        y1 = H_eff if y_end is None else min(y_end, H_eff)
        if self.lib.rvpt_spirv_vm_dispatch(self.vm, x_begin, x1, y_begin, y1, self.nthreads):
            self._check()

    @property
    def executed(self) -> int:
        return int(self.lib.rvpt_spirv_vm_executed(self.vm))

    def close(self):
        if self.vm:
            self.lib.rvpt_spirv_vm_destroy(self.vm)
            self.vm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
