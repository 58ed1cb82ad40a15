"""Converts the reference's built-in model into the engine's asset format.

Run in the build container only (needs /root/reference):

    python tools/make_builtin_scene.py

Reads /root/reference/assets/models/rabbit.obj with the same conventions as
`load_model()` (src/rvpt/main.cpp:12-62 — tinyobjloader, triangulated faces,
positions only) and writes rvpt_b200/assets/builtin_bunny.npz:
    vertices  float32 [V, 3]   the `v` records, parsed to float32
    faces     int32   [F, 3]   zero-based vertex indices per triangle
The GPU box has no /root/reference, so the asset is committed.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from rvpt_b200.scene import parse_obj  # noqa: E402


def main() -> None:
    src = Path("/root/reference/assets/models/rabbit.obj")
    vertices, faces = parse_obj(src.read_text())
    out = ROOT / "rvpt_b200" / "assets" / "builtin_bunny.npz"
    np.savez_compressed(out, vertices=vertices, faces=faces)
    print(f"{out}: {len(vertices)} vertices, {len(faces)} triangles")


if __name__ == "__main__":
    main()
