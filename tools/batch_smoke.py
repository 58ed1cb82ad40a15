"""A small batched launch (rvpt_b200_render_frames) checked against the oracle — the workload
tools/gpu_sanitize.sh runs under compute-sanitizer."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
import rvpt_b200 as rv  # noqa: E402

for scene, pose, fov in ((rv.builtin_scene(), (0.0, 0.0, 0.0), 90.0), (rv.cornell_scene(), (0.0, 1.2, -3.4), 60.0)):
    W, H, N = 96, 64, 5
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = np.ascontiguousarray(scene.triangles[perm])
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H)
    eng.upload_scene(tris, scene.materials, nodes)
    ora = oracle.OracleRenderer(W, H, tris, scene.materials, nodes)
    for start in (0, N):
        eng.render_frames(rv.default_settings(frame=start), cam, N)
        for f in range(start, start + N):
            ora.render_frame(rv.default_settings(frame=f), cam)
    assert np.array_equal(eng.read_accum_f32().view(np.uint32), ora.accum.view(np.uint32)), scene.name
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    print("batch ok:", scene.name, eng.stats())
    eng.close()
