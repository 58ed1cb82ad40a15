"""Python host mirror of the reference's renderer facade for the compute path.

`Engine` plays the role `class RVPT` plays for the Vulkan path
(src/rvpt/rvpt.h:27-90): scene in, `update()`/`draw()`-style frames out — but
every frame goes through the C ABI of include/rvpt_abi.h into the sm_100a
kernels. Nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .scene import BVH_NODE_DTYPE, MATERIAL_DTYPE, RENDER_SETTINGS_DTYPE, TRIANGLE_DTYPE, Scene


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rvpt_b200 error {code}: {message}")
        self.code = code


def build_bvh(triangles: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """BinnedBvhBuilder::build_bvh counterpart (host only). Returns
    (nodes BVH_NODE_DTYPE[], primitive_indices uint32[]); upload
    `triangles[primitive_indices]` like `permute_primitives` (bvh.h:70-77)."""
    lib = _lib.load()
    tris = np.ascontiguousarray(triangles, TRIANGLE_DTYPE)
    n = len(tris)
    nodes = np.zeros(max(2 * n, 1), BVH_NODE_DTYPE)
    perm = np.zeros(n, np.uint32)
    n_nodes = C.c_size_t(0)
    rc = lib.rvpt_b200_build_bvh(tris.ctypes.data, n, nodes.ctypes.data, C.byref(n_nodes),
                                 perm.ctypes.data)
    if rc:
        raise EngineError(rc, "build_bvh failed")
    return nodes[: n_nodes.value].copy(), perm


def build_bvh_gpu(triangles: np.ndarray, device: int = 0) -> tuple[np.ndarray, np.ndarray, float]:
    """The same contract built on the GPU (linear BVH, rvpt_b200/csrc/bvh_gpu.cu). Returns
    (nodes, primitive_indices, device milliseconds of the build kernels)."""
    lib = _lib.load()
    tris = np.ascontiguousarray(triangles, TRIANGLE_DTYPE)
    n = len(tris)
    nodes = np.zeros(max(2 * n, 1), BVH_NODE_DTYPE)
    perm = np.zeros(n, np.uint32)
    n_nodes = C.c_size_t(0)
    ms = C.c_float(0.0)
    rc = lib.rvpt_b200_build_bvh_gpu(device, tris.ctypes.data, n, nodes.ctypes.data, C.byref(n_nodes),
                                     perm.ctypes.data, C.byref(ms))
    if rc:
        raise EngineError(rc, "build_bvh_gpu failed")
    return nodes[: n_nodes.value].copy(), perm, float(ms.value)


def camera_data(translation=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0), aspect: float = 2.0,
                fov: float = 90.0, scale: float = 4.0) -> np.ndarray:
    """Camera::get_data() (camera.cpp:55-66): 20 floats. Defaults are the
    reference's (camera.h:44-49; aspect = window width / height)."""
    lib = _lib.load()
    t = np.asarray(translation, np.float32)
    r = np.asarray(rotation, np.float32)
    out = np.zeros(20, np.float32)
    lib.rvpt_b200_camera_data(t.ctypes.data, r.ctypes.data, float(aspect), float(fov), float(scale),
                              out.ctypes.data)
    return out


class Engine:
    """One rendering context on one GPU (one process per GPU)."""

    def __init__(self, width: int, height: int, device: int = 0, flags: int = 0,
                 rank: int = 0, nranks: int = 1):
        self._lib = _lib.load()
        self.width, self.height = int(width), int(height)
        # RVPT_B200_EXTRA_FLAGS: developer knob to A/B kernel variants under bench.py / tools
        flags |= int(os.environ.get("RVPT_B200_EXTRA_FLAGS", "0"), 0)
        self.flags = flags
        self._ctx = C.c_void_p()
        rc = self._lib.rvpt_b200_create(C.byref(self._ctx), device, width, height, flags)
        if rc:
            msg = self._lib.rvpt_b200_last_error(self._ctx).decode() if self._ctx else "create failed"
            self._lib.rvpt_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()
            raise EngineError(rc, msg)
        if nranks != 1:
            self._check(self._lib.rvpt_b200_set_partition(self._ctx, rank, nranks))
        self.rank, self.nranks = rank, nranks

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc:
            raise EngineError(rc, self._lib.rvpt_b200_last_error(self._ctx).decode())

    def close(self) -> None:
        if self._ctx:
            self._lib.rvpt_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- scene ------------------------------------------------------------
    def upload_scene(self, triangles: np.ndarray, materials: np.ndarray,
                     nodes: np.ndarray | None = None) -> None:
        """rvpt.cpp:123-126. `triangles` in BVH-permuted order when `nodes`
        is given; `nodes=None` builds the BVH inside the library."""
        tris = np.ascontiguousarray(triangles, TRIANGLE_DTYPE)
        mats = np.ascontiguousarray(materials, MATERIAL_DTYPE)
        if nodes is None:
            self._check(self._lib.rvpt_b200_upload_scene(self._ctx, None, 0, tris.ctypes.data,
                                                         len(tris), mats.ctypes.data, len(mats)))
        else:
            nd = np.ascontiguousarray(nodes, BVH_NODE_DTYPE)
            self._check(self._lib.rvpt_b200_upload_scene(self._ctx, nd.ctypes.data, len(nd),
                                                         tris.ctypes.data, len(tris),
                                                         mats.ctypes.data, len(mats)))

    def upload(self, scene: Scene) -> tuple[np.ndarray, np.ndarray]:
        """Builds the BVH on the host, permutes and uploads — what
        RVPT::initialize() does (rvpt.cpp:84-86). Returns (nodes, sorted
        triangles) so a checker can be fed the same bytes."""
        nodes, perm = build_bvh(scene.triangles)
        sorted_tris = scene.triangles[perm]
        self.upload_scene(sorted_tris, scene.materials, nodes)
        return nodes, sorted_tris

    # -- frames -----------------------------------------------------------
    def render_frame(self, settings: np.ndarray, camera: np.ndarray) -> None:
        rs = np.ascontiguousarray(settings, RENDER_SETTINGS_DTYPE)
        cam = np.ascontiguousarray(camera, np.float32)
        assert cam.size == 20
        self._check(self._lib.rvpt_b200_render_frame(self._ctx, rs.ctypes.data, cam.ctypes.data))

    def render_frames(self, settings: np.ndarray, camera: np.ndarray, n_frames: int) -> None:
        """n_frames progressive frames starting at settings.current_frame, one launch."""
        rs = np.ascontiguousarray(settings, RENDER_SETTINGS_DTYPE)
        cam = np.ascontiguousarray(camera, np.float32)
        assert cam.size == 20
        self._check(self._lib.rvpt_b200_render_frames(self._ctx, rs.ctypes.data, cam.ctypes.data,
                                                      int(n_frames)))

    def render_frame_raw(self, settings_ptr: int, camera_ptr: int, n_frames: int = 1) -> None:
        self._check(self._lib.rvpt_b200_render_frames(self._ctx, settings_ptr, camera_ptr, n_frames))

    def sync(self) -> None:
        self._check(self._lib.rvpt_b200_sync(self._ctx))

    def set_stream(self, cuda_stream: int | None) -> None:
        self._check(self._lib.rvpt_b200_set_stream(self._ctx, cuda_stream))

    def reset_accum(self) -> None:
        self._check(self._lib.rvpt_b200_reset_accum(self._ctx))

    # -- read-back --------------------------------------------------------
    def read_output_rgba8(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        self._check(self._lib.rvpt_b200_read_output_rgba8(self._ctx, out.ctypes.data))
        return out

    def read_output_rgba8_async(self, out: np.ndarray) -> None:
        """Enqueues the read-back of the last frames' image into `out` (pinned host memory, HxWx4
        uint8) and returns; later frames fill the second image. Pair with wait_output()."""
        assert out.nbytes == self.height * self.width * 4
        self._check(self._lib.rvpt_b200_read_output_rgba8_async(self._ctx, out.ctypes.data))

    def wait_output(self) -> None:
        self._check(self._lib.rvpt_b200_wait_output(self._ctx))

    def flip_output(self) -> None:
        self._check(self._lib.rvpt_b200_flip_output(self._ctx))

    def read_accum_f32(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), np.float32)
        self._check(self._lib.rvpt_b200_read_accum_f32(self._ctx, out.ctypes.data))
        return out

    def write_accum_f32(self, accum: np.ndarray) -> None:
        a = np.ascontiguousarray(accum, np.float32)
        assert a.shape == (self.height, self.width, 4)
        self._check(self._lib.rvpt_b200_write_accum_f32(self._ctx, a.ctypes.data))

    def stats(self) -> dict:
        st = _lib.Stats()
        self._check(self._lib.rvpt_b200_get_stats(self._ctx, C.byref(st)))
        active = [int(v) for v in st.active]
        while active and active[-1] == 0:
            active.pop()
        return {"samples": int(st.samples), "segments": int(st.segments), "active": active,
                "kernel_launches": int(st.kernel_launches), "traversal_order": int(st.traversal_order),
                "frames": int(st.frames)}

    def set_profiling(self, enabled: bool) -> None:
        self._check(self._lib.rvpt_b200_set_profiling(self._ctx, int(enabled)))

    def kernel_times(self) -> dict:
        kt = _lib.KernelTimes()
        self._check(self._lib.rvpt_b200_get_kernel_times(self._ctx, C.byref(kt)))
        return {"primary_ms": kt.primary_ms, "bounce_ms": kt.bounce_ms,
                "primary_launches": kt.primary_launches, "bounce_launches": kt.bounce_launches}

    def set_timeline(self, enabled: bool) -> None:
        """Per-CTA phase stamps of the frame kernel (after upload_scene)."""
        self._check(self._lib.rvpt_b200_set_timeline(self._ctx, int(enabled)))

    def timeline(self) -> np.ndarray:
        """uint64 [n_ctas, n_slots] %globaltimer stamps (ns) of the last frame kernel, 0 =
        phase not reached; slot meaning in include/rvpt_abi.h."""
        n_ctas, n_slots = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.rvpt_b200_get_timeline(self._ctx, None, 0, C.byref(n_ctas),
                                                     C.byref(n_slots)))
        out = np.zeros((n_ctas.value, n_slots.value), np.uint64)
        if out.size:
            self._check(self._lib.rvpt_b200_get_timeline(self._ctx, out.ctypes.data, out.size,
                                                         C.byref(n_ctas), C.byref(n_slots)))
        return out

    # -- multi-GPU tiles --------------------------------------------------
    def tile_info(self) -> _lib.TileInfo:
        ti = _lib.TileInfo()
        self._check(self._lib.rvpt_b200_get_tile_info(self._ctx, C.byref(ti)))
        return ti

    def set_external_tiles(self, d_accum: int | None, d_rgba8: int | None) -> None:
        self._check(self._lib.rvpt_b200_set_external_tiles(self._ctx, d_accum, d_rgba8))

    def export_output(self, second: bool = False) -> bytes:
        """Display rank: CUDA IPC handle of the raster result image (or of its double buffer)."""
        buf = (C.c_ubyte * 64)()
        fn = self._lib.rvpt_b200_export_output2 if second else self._lib.rvpt_b200_export_output
        self._check(fn(self._ctx, buf))
        return bytes(buf)

    def attach_output(self, handle: bytes, second: bool = False) -> None:
        """Other ranks: write finished pixels straight into the display rank's image."""
        assert len(handle) == 64
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        fn = self._lib.rvpt_b200_attach_output2 if second else self._lib.rvpt_b200_attach_output
        self._check(fn(self._ctx, buf))

    def untile(self, d_gathered: int, d_raster: int, elem_bytes: int,
               cuda_stream: int | None = None) -> None:
        if cuda_stream is None:
            self._check(self._lib.rvpt_b200_untile(self._ctx, d_gathered, d_raster, elem_bytes,
                                                   self.nranks))
        else:
            self._check(self._lib.rvpt_b200_untile_on(self._ctx, d_gathered, d_raster, elem_bytes,
                                                      self.nranks, cuda_stream))
