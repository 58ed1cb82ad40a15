"""Writes oracle/workloads/*.npz: the byte-exact inputs of the bench workloads (BVH nodes,
BVH-ordered triangles, materials, the 80-byte camera block) so that bench.py's reference arm
and its parity check can run WITHOUT loading the product library. Generated here with the
product's own host helpers (rvpt_b200.build_bvh / camera_data); bench.py's GPU arm asserts that
what it builds at run time equals these files."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import rvpt_b200 as rv  # noqa: E402

OUT = ROOT / "oracle" / "workloads"
WORKLOADS = {
    # name: (scene factory, pose, fov)
    "builtin_default": (rv.builtin_scene, (0.0, 0.0, 0.0), 90.0),
    "builtin_pinned": (rv.builtin_scene, (0.0, 0.8, -2.5), 90.0),
    "cornell_default": (rv.cornell_scene, (0.0, 1.2, -3.4), 60.0),
}

for name, (factory, pose, fov) in WORKLOADS.items():
    scene = factory()
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = np.ascontiguousarray(scene.triangles[perm])
    cam = rv.camera_data(translation=pose, aspect=16 / 9, fov=fov)  # 1920x1080 and 3840x2160
    np.savez_compressed(OUT / f"{name}.npz", nodes=nodes, triangles=tris, materials=scene.materials,
                        camera_16x9=cam, pose=np.asarray(pose, np.float32), fov=np.float32(fov))
    print(name, len(tris), "triangles", len(nodes), "nodes", (OUT / f"{name}.npz").stat().st_size, "bytes")
