"""Where the time goes inside the persistent frame kernel: per-CTA %globaltimer stamps
(rvpt_b200_set_timeline) summarised per phase. Usage (GPU box):
    python tools/timeline.py [--scene builtin|cornell] [--pose default|pinned] [--flags N] > profiles/...md
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rvpt_b200 as rv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="builtin")
ap.add_argument("--pose", default="default")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--flags", type=lambda v: int(v, 0), default=0)
ap.add_argument("--frames", type=int, default=12)
ap.add_argument("--batch", type=int, default=0, help="frames per batched launch (0: one launch per frame)")
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--nranks", type=int, default=1, help="render one rank's tile set of an N-way partition")
args = ap.parse_args()

scene = rv.builtin_scene() if args.scene == "builtin" else rv.cornell_scene()
pose = {"default": (0.0, 0.0, 0.0), "pinned": (0.0, 0.8, -2.5)}[args.pose]
fov = 90.0
if args.scene == "cornell":
    pose, fov = (0.0, 1.2, -3.4), 60.0
W, H = args.width, args.height
cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
eng = rv.Engine(W, H, flags=args.flags, rank=args.rank, nranks=args.nranks)
eng.upload(scene)
eng.render_frames(rv.default_settings(frame=0), cam, max(4, args.batch))
eng.sync()
eng.set_timeline(True)
rows = []
for f in range(4, 4 + args.frames):
    if args.batch:
        eng.render_frames(rv.default_settings(frame=0), cam, args.batch)
    else:
        eng.render_frame(rv.default_settings(frame=f), cam)
    tl = eng.timeline().astype(np.int64)
    t0 = tl[:, 0].min()
    rel = np.where(tl > 0, tl - t0, -1) / 1e3  # us
    rows.append(rel)
st = eng.stats()
print(f"# frame-kernel timeline, {args.scene} scene, {W}x{H}, pose {pose}, flags {args.flags:#x}, "
      f"{'one launch per frame' if not args.batch else str(args.batch) + ' frames per launch'}, "
      f"rank {args.rank} of {args.nranks}")
print(f"active rays per bounce (last launch, {st['frames']} frame(s)): {st['active']}")
print()
print("Per phase stamp, over CTAs (us since the first CTA entered the kernel), median over "
      f"{args.frames} frames of (min / median / max over CTAs):")
print()
print("| slot | meaning | min | median | max |")
print("|---|---|---|---|---|")
names = {0: "kernel entry", 1: "scene staged + copies derived", 2: "primary wave done (warp 0)"}
for b in range(1, 7):
    names[2 * b + 1] = f"past grid barrier before wave {b}"
    names[2 * b + 2] = f"wave {b} done (warp 0)"
if args.batch:
    names[14], names[15] = "every sample parked (past the barrier before the resolve)", "resolve phase done (warp 0)"
for k in range(16):
    vals = [(r[:, k][r[:, k] >= 0]) for r in rows]
    if not all(len(v) for v in vals):
        continue
    mn = np.median([v.min() for v in vals])
    md = np.median([np.median(v) for v in vals])
    mx = np.median([v.max() for v in vals])
    print(f"| {k} | {names.get(k, '')} | {mn:.1f} | {md:.1f} | {mx:.1f} |")
