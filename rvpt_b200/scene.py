"""Host-side scene construction: the near side of the drop-in boundary.

Mirrors what `main.cpp` does before `RVPT::initialize()` — load an OBJ into
`Triangle`s, add `Material`s (src/rvpt/main.cpp:12-62, 102-107) — and produces
numpy arrays whose bytes are exactly the reference's PODs (include/rvpt_abi.h).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np

ASSETS = Path(__file__).resolve().parent / "assets"

# Material::Type, src/rvpt/material.h:11-16
LAMBERT, MIRROR, DIELECTRIC = 0, 1, 2

TRIANGLE_DTYPE = np.dtype(
    [("vertex0", "<f4", 4), ("vertex1", "<f4", 4), ("vertex2", "<f4", 4), ("material_id", "<f4", 4)]
)
MATERIAL_DTYPE = np.dtype([("albedo", "<f4", 4), ("emission", "<f4", 4), ("data", "<f4", 4)])
BVH_NODE_DTYPE = np.dtype(
    [("first_child_or_primitive", "<u4"), ("primitive_count", "<u4"), ("bounds", "<f4", 6)]
)
RENDER_SETTINGS_DTYPE = np.dtype(
    [
        ("max_bounces", "<i4"),
        ("aa", "<i4"),
        ("current_frame", "<u4"),
        ("camera_mode", "<i4"),
        ("top_left_render_mode", "<i4"),
        ("top_right_render_mode", "<i4"),
        ("bottom_left_render_mode", "<i4"),
        ("bottom_right_render_mode", "<i4"),
        ("split_ratio", "<f4", 2),
    ]
)
assert TRIANGLE_DTYPE.itemsize == 64 and MATERIAL_DTYPE.itemsize == 48
assert BVH_NODE_DTYPE.itemsize == 32 and RENDER_SETTINGS_DTYPE.itemsize == 40


def parse_obj(text: str) -> tuple[np.ndarray, np.ndarray]:
    """Minimal OBJ reader with tinyobjloader's conventions for what
    `load_model` consumes: `v x y z` positions, `f` records whose indices may
    be `i`, `i/t`, `i//n` or `i/t/n`, 1-based or negative (relative), polygons
    fan-triangulated. Returns (vertices float32 [V,3], faces int32 [F,3])."""
    verts: list[tuple[float, float, float]] = []
    faces: list[tuple[int, int, int]] = []
    for line in text.splitlines():
        parts = line.split()
        if not parts:
            continue
        if parts[0] == "v":
            verts.append((float(parts[1]), float(parts[2]), float(parts[3])))
        elif parts[0] == "f":
            idx = []
            for tok in parts[1:]:
                i = int(tok.split("/")[0])
                idx.append(i - 1 if i > 0 else len(verts) + i)
            for k in range(1, len(idx) - 1):
                faces.append((idx[0], idx[k], idx[k + 1]))
    return (np.asarray(verts, dtype=np.float32).reshape(-1, 3),
            np.asarray(faces, dtype=np.int32).reshape(-1, 3))


def write_obj(path, vertices: np.ndarray, faces: np.ndarray) -> None:
    """Writes positions + triangles as a Wavefront OBJ (9 significant digits, so
    parsing it back reproduces the float32 values)."""
    with open(path, "w") as f:
        f.write("o mesh\n")
        for v in np.asarray(vertices, np.float32):
            f.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for t in np.asarray(faces):
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def make_triangles(v0: np.ndarray, v1: np.ndarray, v2: np.ndarray, material_id) -> np.ndarray:
    """`Triangle(v0, v1, v2, material_id)` (src/rvpt/geometry.h:81-91): the
    face normal normalize(cross(v1-v0, v2-v0)) is packed into the three .w."""
    v0 = np.asarray(v0, np.float32).reshape(-1, 3)
    v1 = np.asarray(v1, np.float32).reshape(-1, 3)
    v2 = np.asarray(v2, np.float32).reshape(-1, 3)
    n = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    length = np.sqrt((n * n).sum(axis=1, keepdims=True, dtype=np.float32))
    with np.errstate(invalid="ignore", divide="ignore"):
        n = (n / length).astype(np.float32)
    tris = np.zeros(len(v0), TRIANGLE_DTYPE)
    tris["vertex0"][:, :3] = v0
    tris["vertex1"][:, :3] = v1
    tris["vertex2"][:, :3] = v2
    tris["vertex0"][:, 3] = n[:, 0]
    tris["vertex1"][:, 3] = n[:, 1]
    tris["vertex2"][:, 3] = n[:, 2]
    tris["material_id"][:, 0] = np.asarray(material_id, np.float32)
    return tris


def triangles_from_mesh(vertices: np.ndarray, faces: np.ndarray, material_id) -> np.ndarray:
    return make_triangles(vertices[faces[:, 0]], vertices[faces[:, 1]], vertices[faces[:, 2]],
                          material_id)


def make_material(albedo, emission=(0, 0, 0, 0), mtype: int = LAMBERT) -> np.ndarray:
    """`Material(albedo, emission, type)` (src/rvpt/material.h:17-22).
    albedo.w doubles as the index of refraction (intersection.glsl:54)."""
    m = np.zeros(1, MATERIAL_DTYPE)
    a = list(albedo) + [0.0] * (4 - len(albedo))
    e = list(emission) + [0.0] * (4 - len(emission))
    m["albedo"][0] = np.asarray(a, np.float32)
    m["emission"][0] = np.asarray(e, np.float32)
    m["data"][0, 0] = float(mtype)
    return m


@dataclass
class Scene:
    triangles: np.ndarray  # TRIANGLE_DTYPE, upload order
    materials: np.ndarray  # MATERIAL_DTYPE
    name: str = "scene"


def load_obj(path: str | Path, material_id: int = 1) -> np.ndarray:
    """`load_model(rvpt, path, material_id)` (src/rvpt/main.cpp:12-62)."""
    vertices, faces = parse_obj(Path(path).read_text())
    return triangles_from_mesh(vertices, faces, material_id)


def builtin_mesh() -> tuple[np.ndarray, np.ndarray]:
    data = np.load(ASSETS / "builtin_bunny.npz")
    return data["vertices"], data["faces"]


def builtin_scene() -> Scene:
    """The scene `main()` builds (src/rvpt/main.cpp:102-107): rabbit.obj with
    material 1; material 0 = white Lambert with emission (0.1,0.4,0.6) (unused
    by any triangle), material 1 = white Lambert, no emission."""
    vertices, faces = builtin_mesh()
    tris = triangles_from_mesh(vertices, faces, 1)
    mats = np.concatenate([
        make_material((1, 1, 1, 0), (0.1, 0.4, 0.6, 0), LAMBERT),
        make_material((1.0, 1.0, 1.0, 0), (0, 0, 0, 0), LAMBERT),
    ])
    return Scene(tris, mats, "builtin")


def _quad(a, b, c, d, material_id) -> np.ndarray:
    a, b, c, d = (np.asarray(p, np.float32) for p in (a, b, c, d))
    return make_triangles(np.stack([a, a]), np.stack([b, c]), np.stack([c, d]), material_id)


def _box(lo, hi, material_id) -> np.ndarray:
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    p = lambda x, y, z: (x, y, z)  # noqa: E731
    return np.concatenate([
        _quad(p(x0, y0, z0), p(x0, y1, z0), p(x1, y1, z0), p(x1, y0, z0), material_id),  # -z
        _quad(p(x0, y0, z1), p(x1, y0, z1), p(x1, y1, z1), p(x0, y1, z1), material_id),  # +z
        _quad(p(x0, y0, z0), p(x0, y0, z1), p(x0, y1, z1), p(x0, y1, z0), material_id),  # -x
        _quad(p(x1, y0, z0), p(x1, y1, z0), p(x1, y1, z1), p(x1, y0, z1), material_id),  # +x
        _quad(p(x0, y0, z0), p(x1, y0, z0), p(x1, y0, z1), p(x0, y0, z1), material_id),  # -y
        _quad(p(x0, y1, z0), p(x0, y1, z1), p(x1, y1, z1), p(x1, y1, z0), material_id),  # +y
    ])


def cornell_scene(with_bunny: bool = True, with_blocks: bool = True, block_gap: float = 0.01) -> Scene:
    """BASELINE.json config 3: a Cornell box generated in code, open towards
    the camera (a closed box renders black: paths that exhaust their bounces
    return 0, integrators.glsl:674-675), an emissive ceiling patch, red / green
    / white Lambert walls, optional mirror + dielectric blocks, the bunny
    inside. Camera: position (0, 1, -3.2) looking +z.

    The blocks hover `block_gap` above the floor (as the light patch hangs 1 cm
    below the ceiling) so that no two faces of the scene coincide; with
    `block_gap=0` their bottom faces lie in the floor plane, every ray through
    them ties in t, and the engine keeps the reference's BVH child order
    (rvpt_abi.h, RVPT_B200_FLAG_REFERENCE_ORDER) — the tests use both."""
    W, R, G, LIGHT, MIRR, GLASS = 0, 1, 2, 3, 4, 5
    mats = np.concatenate([
        make_material((0.73, 0.73, 0.73, 0)),
        make_material((0.65, 0.05, 0.05, 0)),
        make_material((0.12, 0.45, 0.15, 0)),
        make_material((0.78, 0.78, 0.78, 0), (15, 15, 15, 0)),
        make_material((0.9, 0.9, 0.9, 0), mtype=MIRROR),
        make_material((1.0, 1.0, 1.0, 1.5), mtype=DIELECTRIC),
    ])
    x0, x1, y0, y1, z0, z1 = -1.2, 1.2, 0.0, 2.4, -1.2, 1.2
    parts = [
        _quad((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1), W),   # floor
        _quad((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0), W),   # ceiling
        _quad((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1), W),   # back wall
        _quad((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0), R),   # left
        _quad((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1), G),   # right
        _quad((-0.4, y1 - 0.01, -0.4), (-0.4, y1 - 0.01, 0.4), (0.4, y1 - 0.01, 0.4),
              (0.4, y1 - 0.01, -0.4), LIGHT),
    ]
    if with_blocks:
        g = float(block_gap)
        parts.append(_box((-0.95, g, 0.35), (-0.35, 1.3 + g, 0.95), MIRR))
        parts.append(_box((0.45, g, -0.65), (0.95, 0.6 + g, -0.15), GLASS))
    if with_bunny:
        vertices, faces = builtin_mesh()
        v = vertices * np.float32(0.55) + np.asarray([0.15, 0.0, 0.1], np.float32)
        parts.append(triangles_from_mesh(v.astype(np.float32), faces, W))
    return Scene(np.concatenate(parts), mats, "cornell")


def displaced_sphere_scene(n_triangles: int = 20000, seed: int = 1) -> Scene:
    """A procedurally generated large mesh (SURVEY.md §8 f-1 stand-in for the
    reference's 560 k-triangle tridel-interior-test.obj, which is not loaded by
    main.cpp and cannot travel to the GPU box): a UV sphere of radius ~1 at
    (0, 1, 0) with smooth radial displacement, ~n_triangles triangles, three
    Lambert materials in latitude bands, over a large ground quad. Scenes whose
    device blob (plus its derived copies) exceeds the 224 KB shared-memory budget take the L2-resident (non-shared-memory) kernel path."""
    nv = max(4, int(round((n_triangles / 4.0) ** 0.5)))
    nu = 2 * nv
    u = np.linspace(0.0, 2.0 * np.pi, nu + 1)[:-1]
    v = np.linspace(0.0, np.pi, nv + 1)
    uu, vv = np.meshgrid(u, v)                      # [nv+1, nu]
    rng = np.random.default_rng(seed)
    k = rng.uniform(1.0, 6.0, size=(6, 2))
    ph = rng.uniform(0.0, 2.0 * np.pi, size=6)
    r = 1.0 + sum(0.035 * np.sin(k[i, 0] * uu + ph[i]) * np.sin(k[i, 1] * vv) for i in range(6))
    x = r * np.sin(vv) * np.cos(uu)
    y = 1.0 + r * np.cos(vv)
    z = r * np.sin(vv) * np.sin(uu)
    pts = np.stack([x, y, z], axis=-1).astype(np.float32)   # [nv+1, nu, 3]
    i0 = np.arange(nv)[:, None]
    j0 = np.arange(nu)[None, :]
    j1 = (j0 + 1) % nu
    a, b, c, d = pts[i0, j0], pts[i0 + 1, j0], pts[i0 + 1, j1], pts[i0, j1]
    band = ((i0 * 3) // nv + 0 * j0).astype(np.float32)      # material 0..2 by latitude
    t1 = make_triangles(a.reshape(-1, 3), b.reshape(-1, 3), c.reshape(-1, 3), band.reshape(-1))
    t2 = make_triangles(a.reshape(-1, 3), c.reshape(-1, 3), d.reshape(-1, 3), band.reshape(-1))
    tris = np.concatenate([t1, t2])
    # drop the degenerate triangles at the poles
    e0 = tris["vertex1"][:, :3] - tris["vertex0"][:, :3]
    e1 = tris["vertex2"][:, :3] - tris["vertex0"][:, :3]
    area2 = np.linalg.norm(np.cross(e0, e1), axis=1)
    tris = tris[area2 > 1e-12]
    ground = _quad((-30, -0.2, -30), (-30, -0.2, 30), (30, -0.2, 30), (30, -0.2, -30), 3)
    mats = np.concatenate([
        make_material((0.75, 0.25, 0.2, 0)), make_material((0.8, 0.8, 0.8, 0)),
        make_material((0.2, 0.35, 0.75, 0)), make_material((0.5, 0.5, 0.5, 0)),
    ])
    return Scene(np.concatenate([tris, ground]), mats, f"displaced_sphere_{len(tris) + 2}")


TRIDEL_NPZ = ASSETS / "tridel_interior.npz"


def tridel_scene() -> tuple[Scene, tuple[float, float, float], float]:
    """The reference's large test model, assets/models/tridel-interior-test.obj (560 021
    triangles), from the packed copy tools/pack_tridel.py writes next to the
    reference tree (rvpt_b200/assets/tridel_interior.npz: git-ignored, travels with the repo
    snapshot). One white Lambert material like `load_model(rvpt, path, material_id)`
    (main.cpp:12-62) plus the unused emissive material 0 of main.cpp:102-107; the degenerate
    triangles of the file (1 318 with zero area) are kept — the reference uploads them too.
    Returns (scene, camera pose inside the room, fov)."""
    if not TRIDEL_NPZ.exists():
        raise FileNotFoundError(f"{TRIDEL_NPZ} is missing: run tools/pack_tridel.py next to the reference tree")
    data = np.load(TRIDEL_NPZ)
    tris = triangles_from_mesh(data["vertices"], data["faces"], 1)
    mats = np.concatenate([
        make_material((1, 1, 1, 0), (0.1, 0.4, 0.6, 0), LAMBERT),
        make_material((0.8, 0.8, 0.8, 0), (0, 0, 0, 0), LAMBERT),
    ])
    return Scene(tris, mats, "tridel_interior"), (-2.5, 1.5, -1.0), 75.0


def default_settings(max_bounces: int = 8, aa: int = 1, frame: int = 0, camera_mode: int = 0,
                     mode: int = 9) -> np.ndarray:
    """`RVPT::RenderSettings` defaults (src/rvpt/rvpt.h:77-89); the first
    rendered frame is 0 (rvpt.cpp:102-107)."""
    rs = np.zeros(1, RENDER_SETTINGS_DTYPE)
    rs["max_bounces"] = max_bounces
    rs["aa"] = aa
    rs["current_frame"] = frame
    rs["camera_mode"] = camera_mode
    for k in ("top_left", "top_right", "bottom_left", "bottom_right"):
        rs[f"{k}_render_mode"] = mode
    rs["split_ratio"] = (0.5, 0.5)
    return rs
