"""Packs the reference's large test model (assets/models/tridel-interior-test.obj, 560 021
triangles, 65 MB of text — never loaded by main.cpp) into
rvpt_b200/assets/tridel_interior.npz: float32 positions + int32 triangle indices, ~7 MB
compressed. The file is git-ignored (it is the reference's dataset re-encoded, like the built-in
bunny) but travels to the GPU box with the repo snapshot; tests and bench.py skip the large-scene
rows when it is absent. Run next to the reference tree:

    python tools/pack_tridel.py
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/assets/models/tridel-interior-test.obj")
OUT = ROOT / "rvpt_b200" / "assets" / "tridel_interior.npz"


def main():
    if not SRC.exists():
        raise SystemExit(f"{SRC} not found")
    t = time.time()
    verts, faces = [], []
    with open(SRC, "rb") as f:
        for line in f:
            if line.startswith(b"v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith(b"f "):
                # tinyobjloader conventions (main.cpp:12-62): i/t/n records, 1-based or negative, fan triangulation
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split(b"/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    faces.append((idx[0], idx[k], idx[k + 1]))
    v = np.asarray(verts, np.float32)
    fa = np.asarray(faces, np.int32)
    np.savez_compressed(OUT, vertices=v, faces=fa)
    print(f"{len(v)} vertices, {len(fa)} triangles, bounds {v.min(0)} .. {v.max(0)}, "
          f"{OUT.stat().st_size / 1e6:.1f} MB, {time.time() - t:.0f} s")


if __name__ == "__main__":
    sys.exit(main())
