"""rvpt_b200 — B200-native path-tracing engine behind RVPT's render-loop ABI.

The package holds only what the hot path needs: `csrc/` (sm_100a kernels + the
C ABI of include/rvpt_abi.h), the ctypes binding, and the host-side mirror of
the reference's scene/camera/settings interface.
"""
from . import _lib  # noqa: F401
from .engine import Engine, EngineError, build_bvh, build_bvh_gpu, camera_data  # noqa: F401
from .scene import (  # noqa: F401
    DIELECTRIC, LAMBERT, MIRROR, Scene, builtin_scene, cornell_scene, default_settings,
    displaced_sphere_scene, load_obj,
    make_material, make_triangles, tridel_scene,
)

__all__ = [
    "Engine", "EngineError", "build_bvh", "build_bvh_gpu", "camera_data", "Scene", "builtin_scene", "cornell_scene",
    "default_settings", "displaced_sphere_scene", "load_obj", "make_material", "make_triangles", "tridel_scene", "LAMBERT", "MIRROR",
    "DIELECTRIC",
]
