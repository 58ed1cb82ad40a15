"""Localises a parity failure: repeated batched launches at full size under flag variants."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle
import rvpt_b200 as rv
from rvpt_b200 import _lib

W, H = 1920, 1080
def prep(sc):
    nodes, perm = rv.build_bvh(sc.triangles)
    return nodes, np.ascontiguousarray(sc.triangles[perm]), sc.materials
cases = [("pinned", prep(rv.builtin_scene()), (0.0, 0.8, -2.5), 90.0, 64),
         ("cornell", prep(rv.cornell_scene()), (0.0, 1.2, -3.4), 60.0, 16)]
for name, (nodes, tris, mats), pose, fov, n in cases:
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    ora = oracle.OracleRenderer(W, H, tris, mats, nodes)
    for f in range(n):
        ora.render_frame(rv.default_settings(frame=f), cam)
    want = ora.accum
    for label, flags in (("default", 0), ("no_forecast", _lib.FLAG_NO_FORECAST), ("no_sort", _lib.FLAG_NO_QUEUE_SORT),
                         ("ref_order", _lib.FLAG_REFERENCE_ORDER), ("no_batch", _lib.FLAG_NO_BATCH)):
        eng = rv.Engine(W, H, flags=flags)
        eng.upload_scene(tris, mats, nodes)
        res = []
        for rep in range(3):
            eng.render_frames(rv.default_settings(frame=0), cam, n)
            got = eng.read_accum_f32()
            bad = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
            res.append(int(bad.sum()))
            if bad.any() and rep == 1:
                ys, xs = np.nonzero(bad)
                print("   first bad pixels", list(zip(xs[:5], ys[:5])), got[ys[0], xs[0]], want[ys[0], xs[0]],
                      "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
        print(name, label, "differing pixels per repetition:", res, eng.stats()["active"][:5])
        eng.close()
