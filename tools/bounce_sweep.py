"""BASELINE config 5: bounce-depth sweep 1..16 at 1920x1080 on one B200 —
Msamples/s, Mrays/s and per-bounce active rays, for the built-in scene and the
Cornell box. Writes profiles/<tag>_bounce_sweep.md (run under gpurun)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rvpt_b200 as rv  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
W, H, FRAMES = 1920, 1080, 64
out = [f"# Bounce-depth sweep (BASELINE config 5), {W}x{H}, aa 1, {FRAMES} progressive frames per point, 1 x B200\n",
       "Device-resident throughput (CUDA events around one rvpt_b200_render_frames call = one batched launch, after a "
       "warm-up call of the same size). "
       "Warp-divergence counters of the 8-bounce point are in the k_frame ncu capture of the same tag "
       "(`smsp__thread_inst_executed_per_inst_executed.ratio`, `smsp__sass_average_branch_targets_threads_uniform.pct`).\n"]
for name, scene, pose, fov in (("built-in scene, default pose", rv.builtin_scene(), (0.0, 0.0, 0.0), 90.0),
                               ("Cornell box (C3)", rv.cornell_scene(), (0.0, 1.2, -3.4), 60.0)):
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = scene.triangles[perm]
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    eng.set_stream(st.cuda_stream)
    eng.upload_scene(tris, scene.materials, nodes)
    out.append(f"\n## {name} ({len(tris)} triangles)\n\n| max_bounces | Msamples/s | Mrays/s | us/frame | rays/sample | active rays per bounce (whole launch) |\n|---|---|---|---|---|---|\n")
    for b in list(range(1, 9)) + [12, 16]:
        eng.render_frames(rv.default_settings(max_bounces=b, frame=0), cam, FRAMES)  # warm-up, same batch size (buffers grow once)
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        eng.render_frames(rv.default_settings(max_bounces=b, frame=0), cam, FRAMES)
        z.record(st)
        torch.cuda.synchronize()
        ms = a.elapsed_time(z) / FRAMES
        stt = eng.stats()
        rps = stt["segments"] / max(stt["samples"], 1)
        msps = W * H / ms / 1e3
        out.append(f"| {b} | {msps:.0f} | {msps * rps:.0f} | {ms * 1e3:.1f} | {rps:.3f} | {stt['active']} |\n")
    eng.close()
(ROOT / "profiles" / f"{tag}_bounce_sweep.md").write_text("".join(out))
print("".join(out))
