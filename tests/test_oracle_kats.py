"""Pins for the CPU oracle. The reference ships no tests / golden vectors for
this path, so these are known-answer tests derived from its
GLSL source: util.glsl:25-50 (hash + xorshift), compute_pass.comp:5-12
(constants), structs.glsl + the shipped SPIR-V decorations (layouts), plus
committed golden images rendered by the oracle itself (regression pins)."""
import ctypes as C
import hashlib
import math
from pathlib import Path

import numpy as np
import pytest

from conftest import DEFAULT_POSE, PINNED_POSE

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_wang_hash_kats(oracle_mod):
    lib = oracle_mod.load()
    kats = {0: 0xC0A9496A, 1: 0x27922C9D, 2: 0xC6793575, 1920: 0xBCCDD74A, 2073599: 0x03EF862A}
    for seed, want in kats.items():
        assert lib.rvpt_oracle_wang_hash(seed) == want, hex(seed)


def _stream(lib, x, y, w, frame, n):
    states = np.zeros(n, np.uint32)
    vals = np.zeros(n, np.float32)
    lib.rvpt_oracle_rand_stream(x, y, w, frame, n, states.ctypes.data, vals.ctypes.data)
    return states, vals


def test_xorshift_stream_kats(oracle_mod):
    lib = oracle_mod.load()
    states, vals = _stream(lib, 0, 0, 1920, 0, 4)
    assert [int(s) for s in states] == [0xD90BC8A8, 0xA3CD8C47, 0x5AE9C9C5, 0x19FA5D8D]
    np.testing.assert_array_equal(
        vals, np.array([0.8478360772, 0.6398551464, 0.3551298380, 0.1014765203], np.float32))
    # consecutive frames only differ by +1 in the seed (util.glsl:36)
    assert int(_stream(lib, 0, 0, 1920, 1, 1)[0][0]) == 0xD90FE889
    assert int(_stream(lib, 1, 0, 1920, 0, 1)[0][0]) == 0x22360E3D
    # p_idx = x + y * width
    a = _stream(lib, 5, 7, 1920, 3, 8)
    b = _stream(lib, 5 + 7 * 1920, 0, 1, 3, 8)
    assert np.array_equal(a[0], b[0])


def test_rand_can_return_exactly_one():
    """float(uint)/2^32 rounds to nearest: states >= 0xFFFFFF80 give 1.0f
     — the oracle keeps that, it does not 'fix' it to [0,1)."""
    assert np.float32(np.uint32(0xFFFFFF80)) / np.float32(4294967296.0) == np.float32(1.0)
    assert np.float32(np.uint32(0xFFFFFF7F)) / np.float32(4294967296.0) == np.float32(0.99999994)


def test_constants_bit_patterns():
    """compute_pass.comp:5-12 literals as float32."""
    bits = lambda v: int(np.float32(v).view(np.uint32))  # noqa: E731
    assert bits(3.1415926535897932384626433832795) == 0x40490FDB
    assert bits(6.283185307179586476925286766559) == 0x40C90FDB
    assert bits(0.31830988618379067153776752674503) == 0x3EA2F983
    assert bits(0.005) == 0x3BA3D70A
    pi, inv_pi = np.float32(3.1415926535897932), np.float32(0.3183098861837907)
    # white Lambert throughput stays exactly 1 on the built-in scene
    assert (np.float32(1.0) * inv_pi) * pi == np.float32(1.0)
    assert (np.float32(0.5) * inv_pi) * pi == np.float32(0.5)


def test_host_code_is_not_fma_contracted(oracle_mod):
    assert oracle_mod.load().rvpt_oracle_contract_probe() == 1


def test_sincos_accuracy_and_symmetry(oracle_mod):
    lib = oracle_mod.load()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 2 * np.pi, 200000), rng.uniform(-100, 100, 20000),
                        np.linspace(0, 2 * np.pi, 4097)]).astype(np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    lib.rvpt_oracle_sincos(x.ctypes.data, len(x), s.ctypes.data, c.ctypes.data)
    xs = x.astype(np.float64)
    # absolute error well inside Vulkan's 2^-11 bound, and ~ulp-level
    assert np.abs(s - np.sin(xs)).max() < 2.5e-7
    assert np.abs(c - np.cos(xs)).max() < 2.5e-7
    z = np.zeros(1, np.float32)
    lib.rvpt_oracle_sincos(z.ctypes.data, 1, s.ctypes.data, c.ctypes.data)
    assert s[0] == 0.0 and c[0] == 1.0
    bad = np.array([np.inf, np.nan, 1e9], np.float32)
    lib.rvpt_oracle_sincos(bad.ctypes.data, 3, s.ctypes.data, c.ctypes.data)
    assert np.isnan(s[:3]).all() and np.isnan(c[:3]).all()


def test_struct_layouts(rv):
    """Sizes/offsets of the PODs (SPIR-V decorations)."""
    from rvpt_b200 import scene as sc
    assert sc.RENDER_SETTINGS_DTYPE.itemsize == 40
    assert [sc.RENDER_SETTINGS_DTYPE.fields[n][1] for n in sc.RENDER_SETTINGS_DTYPE.names] == \
        [0, 4, 8, 12, 16, 20, 24, 28, 32]
    assert sc.TRIANGLE_DTYPE.itemsize == 64
    assert [sc.TRIANGLE_DTYPE.fields[n][1] for n in sc.TRIANGLE_DTYPE.names] == [0, 16, 32, 48]
    assert sc.BVH_NODE_DTYPE.itemsize == 32
    assert [sc.BVH_NODE_DTYPE.fields[n][1] for n in sc.BVH_NODE_DTYPE.names] == [0, 4, 8]
    assert sc.MATERIAL_DTYPE.itemsize == 48
    assert [sc.MATERIAL_DTYPE.fields[n][1] for n in sc.MATERIAL_DTYPE.names] == [0, 16, 32]
    assert rv.camera_data().nbytes == 80


def test_pinhole_camera_ray(oracle_mod, rv):
    """camera.glsl:29-51 on the default camera: centre pixel looks down +z,
    the vertical field is fov, the horizontal one is scaled by aspect."""
    lib = oracle_mod.load()
    cam = rv.camera_data(aspect=2.0, fov=90.0)
    out = np.zeros(6, np.float32)
    lib.rvpt_oracle_camera_ray(cam.ctypes.data, 0, 0.5, 0.5, out.ctypes.data)
    np.testing.assert_array_equal(out[:3], 0)
    np.testing.assert_allclose(out[3:], [0, 0, 1], atol=1e-7)
    lib.rvpt_oracle_camera_ray(cam.ctypes.data, 0, 1.0, 1.0, out.ctypes.data)
    d = out[3:].astype(np.float64)
    np.testing.assert_allclose(d / d[2], [2.0, 1.0, 1.0], rtol=1e-6)  # tan(45deg) = 1, aspect 2
    assert abs(np.linalg.norm(d) - 1) < 1e-6
    # ortho: direction is the camera's z column, origin spans [-scale*aspect, scale*aspect]
    lib.rvpt_oracle_camera_ray(cam.ctypes.data, 1, 1.0, 0.5, out.ctypes.data)
    np.testing.assert_allclose(out, [8, 0, 0, 0, 0, 1], atol=1e-6)


def test_triangle_intersection_semantics(oracle_mod):
    """intersect_triangle_fast (intersection.glsl:267-323): strict bounds on
    t, u, v, u+v; returns the un-normalised normal."""
    lib = oracle_mod.load()
    tri = np.array([0, 0, 2, 1, 0, 2, 0, 1, 2], np.float32)
    out = np.zeros(6, np.float32)

    def hit(o, d, mint=0.0, maxt=np.inf):
        ray = np.array(list(o) + list(d), np.float32)
        h = lib.rvpt_oracle_intersect_triangle(ray.ctypes.data, tri.ctypes.data, mint, maxt,
                                               out.ctypes.data)
        return h, out.copy()

    h, o = hit((0.25, 0.25, 0), (0, 0, 1))
    assert h == 1 and o[0] == 2.0 and tuple(o[1:4]) == (0, 0, 1) and tuple(o[4:6]) == (0.25, 0.25)
    assert hit((0.25, 0.25, 0), (0, 0, 1), maxt=2.0)[0] == 0      # t < maxt is strict
    assert hit((0.25, 0.25, 0), (0, 0, -1))[0] == 0               # behind the origin
    assert hit((0.0, 0.25, 0), (0, 0, 1))[0] == 0                 # u == 0 is outside
    assert hit((0.5, 0.5, 0), (0, 0, 1))[0] == 0                  # u + v == 1 is outside
    assert hit((0.25, 0.25, 0), (1, 0, 0))[0] == 0                # parallel: t = inf/nan fails
    h, o = hit((0.25, 0.25, 0), (0, 0, 2))                        # unnormalised direction: t halves
    assert h == 1 and o[0] == 1.0


def test_aabb_slab_semantics(oracle_mod):
    lib = oracle_mod.load()
    lo = np.array([-1, -1, 1], np.float32)
    hi = np.array([1, 1, 2], np.float32)

    def box(o, d, mint=0.0, maxt=np.inf):
        ray = np.array(list(o) + list(d), np.float32)
        return lib.rvpt_oracle_intersect_aabb(ray.ctypes.data, lo.ctypes.data, hi.ctypes.data,
                                              mint, maxt)

    assert box((0, 0, 0), (0, 0, 1)) == 1          # axis-parallel ray: 1/0 = inf slabs
    assert box((0, 0, 0), (0, 0, 1), maxt=0.5) == 0
    assert box((0, 0, 0), (0, 0, 1), maxt=1.0) == 1  # t1 >= t0 is not strict
    assert box((2, 0, 0), (0, 0, 1)) == 0
    assert box((0, 0, 3), (0, 0, 1)) == 0          # box behind the ray


def test_fresnel_matches_formula(oracle_mod):
    lib = oracle_mod.load()
    ci, co, eta = 0.8, 0.9, 1.0 / 1.5
    rp = (eta * ci - co) / (eta * ci + co)
    rl = (ci - eta * co) / (ci + eta * co)
    assert abs(lib.rvpt_oracle_fresnel(ci, co, eta) - 0.5 * (rp * rp + rl * rl)) < 1e-7
    assert lib.rvpt_oracle_fresnel(1.0, 1.0, 1.0) == 0.0


@pytest.mark.parametrize("pose,counts", [
    (DEFAULT_POSE, [65536, 29872, 6]),
    (PINNED_POSE, [65536, 3963, 345, 64, 12, 5]),
])
def test_c1_ray_counts_and_bvh_independence(oracle_mod, rv, builtin, pose, counts):
    """BASELINE config 1 (256x256, 1 spp, frame 0). The per-bounce ray counts
    equal the independent numpy probe recorded in SURVEY §8(d); BVH traversal
    and list-order brute force give bit-identical images."""
    cam = rv.camera_data(translation=pose, aspect=1.0)
    rs = rv.default_settings()
    bvh = oracle_mod.OracleRenderer(256, 256, builtin.triangles, builtin.materials, builtin.nodes)
    bvh.render_frame(rs, cam)
    assert bvh.active_list() == counts
    brute = oracle_mod.OracleRenderer(256, 256, builtin.triangles, builtin.materials, None)
    brute.render_frame(rs, cam)
    assert np.array_equal(bvh.accum.view(np.uint32), brute.accum.view(np.uint32))
    a = bvh.accum[..., :3]
    if pose == DEFAULT_POSE:
        assert 0.564 < a.min() < 0.565 and 1.399 < a.max() < 1.400
    else:
        # un-normalised sky lookup: negative and > 1 radiance are real
        assert -0.172 < a.min() < -0.170 and 1.342 < a.max() < 1.343


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_images(oracle_mod, rv, builtin, cornell):
    """Committed regression pins (tests/golden/make_golden.py wrote them):
    sha256 of the float32 running mean + a coarse thumbnail."""
    import json
    pins = json.loads((GOLDEN / "oracle_pins.json").read_text())
    from golden.make_golden import CASES, render_case
    for name in CASES:
        acc, res, active = render_case(oracle_mod, rv, {"builtin": builtin, "cornell": cornell}, name)
        pin = pins[name]
        assert active == pin["active"], name
        assert _digest(acc) == pin["accum_sha256"], name
        assert _digest(res) == pin["rgba8_sha256"], name
        thumb = np.load(GOLDEN / f"{name}_thumb.npy")
        h, w = acc.shape[0] // 8, acc.shape[1] // 8
        mine = acc[: h * 8, : w * 8, :3].reshape(h, 8, w, 8, 3).mean(axis=(1, 3))
        np.testing.assert_allclose(mine, thumb, rtol=0, atol=1e-6)


def test_multithreaded_oracle_is_deterministic(oracle_mod, rv, builtin):
    cam = rv.camera_data(translation=PINNED_POSE, aspect=1.5)
    imgs = []
    for nt in (1, 3, 0):
        o = oracle_mod.OracleRenderer(96, 64, builtin.triangles, builtin.materials, builtin.nodes,
                                      nthreads=nt)
        for f in range(3):
            o.render_frame(rv.default_settings(frame=f), cam)
        imgs.append(o.accum.copy())
    assert np.array_equal(imgs[0], imgs[1]) and np.array_equal(imgs[0], imgs[2])


def test_running_mean_and_rgba8_store(oracle_mod, rv, builtin):
    """compute_pass.comp:146-166: frame 0 ignores the previous image; rgba8
    mode quantises the running mean every frame."""
    cam = rv.camera_data(translation=PINNED_POSE, aspect=1.0)
    o = oracle_mod.OracleRenderer(32, 32, builtin.triangles, builtin.materials, builtin.nodes)
    o.accum[:] = 123.0  # stale data must be multiplied by min(frame, 1) = 0
    o.render_frame(rv.default_settings(frame=0), cam)
    first = o.accum.copy()
    assert np.abs(first).max() < 2.0
    o.render_frame(rv.default_settings(frame=1), cam)
    assert not np.array_equal(first, o.accum)
    q = oracle_mod.OracleRenderer(32, 32, builtin.triangles, builtin.materials, builtin.nodes,
                                  flags=oracle_mod.FLAG_ACCUM_RGBA8)
    q.render_frame(rv.default_settings(frame=0), cam)
    want = np.rint(np.clip(first[..., :3], 0, 1) * np.float32(255)).astype(np.uint8)
    assert np.array_equal(q.result[..., :3], want)
    assert not q.result[..., 3].any()  # alpha = 0 (compute_pass.comp:165-166)


def test_hart_sphere_tracer_known_answers(oracle_mod, rv):
    """integrator_Hart (integrators.glsl:681-693) = iterations of the sphere tracer / 31
    (distance_functions.glsl:70-116, MARCH_ITER 32, MARCH_EPS 0.1), worked by hand on one large
    triangle in the plane z = 5 seen by an orthographic camera looking down +z from the origin:
    a ray over the triangle's interior steps the full distance 5 in iteration 0 and stops in
    iteration 1 (radius 0 < 0.1) -> 1/31; index 10, 11 and -1 all select it
    (compute_pass.comp:96-97). A ray that starts 0.05 above a triangle stops in iteration 0 -> 0.
    The same triangle behind the camera: the distance only grows, the march runs out -> 32/31."""
    mats = rv.make_material((1, 1, 1, 0))
    W = H = 8

    def render(z, mode):
        tri = rv.make_triangles([[-50.0, -50.0, z]], [[50.0, -50.0, z]], [[0.0, 80.0, z]], 0)
        nodes, perm = rv.build_bvh(tri)
        o = oracle_mod.OracleRenderer(W, H, tri[perm], mats, nodes)
        # ortho camera (camera.glsl:55-76): parallel rays along the camera's +z, scale 1
        o.render_frame(rv.default_settings(mode=mode, camera_mode=1), rv.camera_data(aspect=1.0, scale=1.0))
        return o.accum[..., :3]

    for mode in (10, 11, -1):
        assert np.array_equal(render(5.0, mode), np.full((H, W, 3), np.float32(1.0) / np.float32(31.0), np.float32))
    assert not render(0.05, 10).any()
    assert np.array_equal(render(-5.0, 10), np.full((H, W, 3), np.float32(32.0) / np.float32(31.0), np.float32))


def test_debug_integrators_sanity(oracle_mod, rv, builtin):
    """integrators.glsl:24-250 on the built-in scene: binary is 0/1 and marks
    exactly the pixels whose depth is finite; color is the material albedo; the
    normal view is 0.5*n+0.5 of a unit vector; Appel is a cosine in [0,1]."""
    W, H = 64, 48
    cam = rv.camera_data(translation=(0.0, 0.8, -2.5), aspect=W / H)

    def render(mode):
        o = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
        o.render_frame(rv.default_settings(mode=mode), cam)
        return o.accum[..., :3].copy()

    binary, color, depth, normal, appel = (render(m) for m in (0, 1, 2, 3, 6))
    hit = binary[..., 0] == 1.0
    assert set(np.unique(binary).tolist()) == {0.0, 1.0} and 0.02 < hit.mean() < 0.2
    assert ((depth[..., 0] > 0) == hit).all()          # 1/(|d| t), t = inf on a miss
    assert (color[hit] == 1.0).all() and not color[~hit].any()   # white Lambert (main.cpp:107)
    n = (normal[hit] - 0.5) / 0.5
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    assert not normal[~hit].any()
    assert (appel[~hit] == 1.0).all() and appel.min() >= 0.0 and appel.max() <= 1.0
