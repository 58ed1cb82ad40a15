"""Per-frame fixed overheads: fused vs unfused, tiny image vs 1080p, bounce limits."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
import rvpt_b200 as rv
from rvpt_b200 import _lib

def timeit(W, H, pose, flags, bounces, frames=200, scene=None):
    scene = scene or rv.builtin_scene()
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = scene.triangles[perm]
    cam = rv.camera_data(translation=pose, aspect=W / H)
    eng = rv.Engine(W, H, flags=flags)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
    eng.upload_scene(tris, scene.materials, nodes)
    rs = [rv.default_settings(max_bounces=bounces, frame=f % 16) for f in range(frames)]
    for f in range(20): eng.render_frame(rs[f], cam)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(st)
    for f in range(frames): eng.render_frame(rs[f], cam)
    b.record(st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / frames * 1e3, (t1 - t0) / frames * 1e6, eng.stats()["active"]

for W, H in ((64, 64), (1920, 1080)):
    for pose in ((0, 0, 0), (0, 0.8, -2.5), (0, 50.0, 0)):
        for flags, name in ((0, "fused"), (_lib.FLAG_UNFUSED, "unfused")):
            for bounces in (1, 2, 3, 8):
                dev, host, act = timeit(W, H, pose, flags, bounces)
                print(f"{W}x{H} pose={pose} {name:8s} bounces={bounces}: {dev:8.1f} us/frame device, {host:6.1f} us/frame host submit, active={act}")
