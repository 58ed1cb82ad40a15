"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle on
the same seeded inputs. The RNG is a pure function of (x, y, W, frame), so
"same seed" = same settings/camera. Bar: BIT-EXACT float32 radiance (tolerance
stated by BASELINE.json is 1e-4 relative; the shared arithmetic contract of
include/rvpt_math.h lets us hold the stronger bar) and identical rgba8 codes.
"""
import numpy as np
import pytest

from conftest import CORNELL_POSE, DEFAULT_POSE, PINNED_POSE

pytestmark = pytest.mark.gpu


def _render_both(rv, oracle_mod, prep, W, H, pose, frames=1, flags=0, oracle_flags=None, fov=90.0,
                 nodes="bvh", **settings_kw):
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H, flags=flags)
    if nodes == "bvh":
        eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    else:
        eng.upload_scene(prep.triangles, prep.materials, None)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes,
                                    flags=flags if oracle_flags is None else oracle_flags)
    stats = []
    for f in range(frames):
        rs = rv.default_settings(frame=f, **settings_kw)
        eng.render_frame(rs, cam)
        ora.render_frame(rs, cam)
        stats.append((eng.stats(), ora.active_list()))
    return eng, ora, stats


def _assert_bit_equal(a, b, what):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    same = a.view(np.uint32) == b.view(np.uint32)
    if not same.all():
        bad = np.argwhere(~same)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
        raise AssertionError(
            f"{what}: {len(bad)} of {same.size} float32 words differ; first at {bad[0]}, "
            f"gpu={a[tuple(bad[0])]!r} oracle={b[tuple(bad[0])]!r}, max rel err {np.nanmax(rel):.3e}")


@pytest.mark.parametrize("pose", [DEFAULT_POSE, PINNED_POSE])
def test_c1_builtin_256_bit_exact(rv, oracle_mod, builtin, pose):
    """BASELINE config 1: built-in scene, 256x256, 1 spp, frame 0."""
    eng, ora, stats = _render_both(rv, oracle_mod, builtin, 256, 256, pose)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "accum")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    st, active = stats[0]
    assert st["active"] == active, "per-bounce ray counts must match the oracle"
    assert st["samples"] == 256 * 256


def test_progressive_frames_and_rel_tolerance(rv, oracle_mod, builtin):
    """8 progressive frames; also states the north-star tolerance explicitly."""
    eng, ora, stats = _render_both(rv, oracle_mod, builtin, 320, 180, (-0.13, 0.83, -1.6), frames=8)
    g, o = eng.read_accum_f32(), ora.accum
    rel = np.abs(g - o) / np.maximum(np.abs(o), 1e-6)
    assert rel.max() <= 1e-4  # BASELINE.json north_star tolerance
    _assert_bit_equal(g, o, "accum after 8 frames")
    for st, active in stats:
        assert st["active"] == active


def test_cornell_mirror_dielectric_bit_exact(rv, oracle_mod, cornell):
    """BASELINE config 3 shape: emissive light, mirror and dielectric blocks."""
    eng, ora, stats = _render_both(rv, oracle_mod, cornell, 200, 152, CORNELL_POSE, frames=3, fov=60.0)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "cornell accum")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    st, active = stats[-1]
    assert st["active"] == active
    assert len(active) == 8 and active[7] > 0, "the Cornell scene must exercise all 8 bounces"


def test_aa_passes_continue_the_rng_stream(rv, oracle_mod, cornell):
    """aa > 1: sample i+1 of a pixel continues the xorshift stream of sample i
    (compute_pass.comp:151-158)."""
    eng, ora, stats = _render_both(rv, oracle_mod, cornell, 96, 80, CORNELL_POSE, frames=2, aa=3, fov=60.0)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "aa=3 accum")
    st, active = stats[-1]
    assert st["active"] == active
    assert st["samples"] == 96 * 80 * 3


@pytest.mark.parametrize("bounces", [1, 2, 16])
def test_bounce_limits(rv, oracle_mod, cornell, bounces):
    eng, ora, stats = _render_both(rv, oracle_mod, cornell, 64, 48, CORNELL_POSE, max_bounces=bounces, fov=60.0)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"max_bounces={bounces}")
    assert stats[0][0]["active"] == stats[0][1]


@pytest.mark.parametrize("camera_mode", [1, 2])
def test_ortho_and_spherical_cameras(rv, oracle_mod, builtin, camera_mode):
    eng, ora, _ = _render_both(rv, oracle_mod, builtin, 128, 64, (0.0, 0.8, -2.5), camera_mode=camera_mode)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"camera_mode={camera_mode}")


def test_rgba8_accumulation_mode(rv, oracle_mod, builtin):
    """Reference-faithful UNORM8 temporal image (rvpt.cpp:759-766): codes must
    be identical (the stated tolerance for a real Vulkan run is 1 LSB)."""
    from rvpt_b200 import _lib
    eng, ora, _ = _render_both(rv, oracle_mod, builtin, 160, 96, PINNED_POSE, frames=5,
                               flags=_lib.FLAG_ACCUM_RGBA8)
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    assert np.array_equal(eng.read_accum_f32(), ora.accum_f32())


def test_reference_dispatch_truncation(rv, oracle_mod, builtin):
    """W/16 x H/16 groups with integer division (rvpt.cpp:1035-1036): the
    remainder rows/columns are never written."""
    from rvpt_b200 import _lib
    W, H = 200, 90  # 12 x 5 groups -> 192 x 80 covered
    eng, ora, stats = _render_both(rv, oracle_mod, builtin, W, H, PINNED_POSE,
                                   flags=_lib.FLAG_REFERENCE_DISPATCH)
    g = eng.read_accum_f32()
    _assert_bit_equal(g, ora.accum, "reference dispatch")
    assert not g[80:].any() and not g[:, 192:].any()
    assert stats[0][0]["samples"] == 192 * 80


def test_brute_force_flag_and_internal_bvh(rv, oracle_mod, builtin):
    """BVH traversal only prunes: list-order nearest hit gives the same image
    on this scene; so does the BVH built inside upload_scene(nodes=NULL)."""
    from rvpt_b200 import _lib
    eng_bvh, ora, _ = _render_both(rv, oracle_mod, builtin, 128, 128, DEFAULT_POSE)
    eng_bf, ora_bf, _ = _render_both(rv, oracle_mod, builtin, 128, 128, DEFAULT_POSE,
                                     flags=_lib.FLAG_BRUTE_FORCE)
    eng_int, _, _ = _render_both(rv, oracle_mod, builtin, 128, 128, DEFAULT_POSE, nodes="internal")
    _assert_bit_equal(eng_bf.read_accum_f32(), ora_bf.accum, "brute force vs oracle brute force")
    _assert_bit_equal(eng_bf.read_accum_f32(), eng_bvh.read_accum_f32(), "brute force vs BVH")
    # internal build permutes the already-permuted triangles again; image is order independent here
    _assert_bit_equal(eng_int.read_accum_f32(), eng_bvh.read_accum_f32(), "internal BVH")


def test_gpu_matches_committed_golden_pins(rv, oracle_mod, builtin, cornell):
    """File-based target: the sha256 pins under tests/golden/ (written by
    tests/golden/make_golden.py from the oracle) — BASELINE configs 1-3 shapes."""
    import hashlib
    import json
    from pathlib import Path
    from golden.make_golden import CASES
    pins = json.loads((Path(__file__).parent / "golden" / "oracle_pins.json").read_text())
    scenes = {"builtin": builtin, "cornell": cornell}
    for name, (scene, W, H, pose, fov, frames, over, flags) in CASES.items():
        prep = scenes[scene]
        cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
        eng = rv.Engine(W, H, flags=flags)
        eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
        for f in range(frames):
            eng.render_frame(rv.default_settings(frame=f, **over), cam)
        acc = eng.read_accum_f32()
        assert hashlib.sha256(acc.tobytes()).hexdigest() == pins[name]["accum_sha256"], name
        assert hashlib.sha256(eng.read_output_rgba8().tobytes()).hexdigest() == \
            pins[name]["rgba8_sha256"], name
        assert eng.stats()["active"] == pins[name]["active"], name
        eng.close()


def test_gpu_reproduces_spirv_pins(rv):
    """The CUDA path against the reference's SHIPPED shader binary: tests/golden/spirv_pins.json
    holds digests of what assets/shaders/compute_pass.comp.spv computes (executed by
    oracle/spirv_vm.cpp; generated by tests/golden/make_spirv_golden.py next to the reference).
    Float32 running mean bit for bit and rgba8 codes, every case: C1 both poses, the reference's
    own UNORM8 + W/16 dispatch configuration, a 1080p band, Cornell, aa, bounce limits, cameras,
    integrators 0-8, split view."""
    import json
    from golden import make_spirv_golden as G
    pins = json.loads(G.PINS.read_text())
    prepared = {}
    for name, s in G.scenes(rv).items():
        nodes, perm = rv.build_bvh(s.triangles)
        prepared[name] = (nodes, np.ascontiguousarray(s.triangles[perm]), s.materials)
    for name, c in G.CASES.items():
        nodes, tris, mats = prepared[c["scene"]]
        flags = (1 if c["unorm8"] else 0) | (2 if c["ref_dispatch"] else 0)
        eng = rv.Engine(c["W"], c["H"], flags=flags)
        eng.upload_scene(tris, mats, nodes)
        cam = rv.camera_data(translation=c["pose"], aspect=c["W"] / c["H"], fov=c["fov"])
        for f in range(c["frames"]):
            eng.render_frame(G.settings_for(rv, c, f), cam)
        got = G.digests(c, eng.read_accum_f32(), eng.read_output_rgba8())
        for key in ("temporal", "rgba8"):
            if got[key]["sha256"] != pins[name][key]["sha256"]:
                rows = [i for i, (a, b) in enumerate(zip(got[key]["rows_crc32"], pins[name][key]["rows_crc32"]))
                        if a != b]
                raise AssertionError(f"{name}/{key}: {len(rows)} rows differ from the SPIR-V run, first {rows[:8]}")
        eng.close()


@pytest.mark.parametrize("aa", [1, 2])
def test_multi_frame_launch_equals_frame_by_frame(rv, oracle_mod, cornell, aa):
    """rvpt_b200_render_frames(n) == n x render_frame with current_frame++ ==
    the oracle, including a batch that does not start at frame 0."""
    W, H = 144, 96
    cam = rv.camera_data(translation=CORNELL_POSE, aspect=W / H, fov=60.0)
    one = rv.Engine(W, H)
    one.upload_scene(cornell.triangles, cornell.materials, cornell.nodes)
    batch = rv.Engine(W, H)
    batch.upload_scene(cornell.triangles, cornell.materials, cornell.nodes)
    ora = oracle_mod.OracleRenderer(W, H, cornell.triangles, cornell.materials, cornell.nodes)
    for f in range(7):
        rs = rv.default_settings(frame=f, aa=aa)
        one.render_frame(rs, cam)
        ora.render_frame(rs, cam)
    batch.render_frames(rv.default_settings(frame=0, aa=aa), cam, 3)
    batch.render_frames(rv.default_settings(frame=3, aa=aa), cam, 4)
    _assert_bit_equal(batch.read_accum_f32(), ora.accum, "multi-frame launch vs oracle")
    _assert_bit_equal(batch.read_accum_f32(), one.read_accum_f32(), "multi-frame vs frame by frame")
    assert np.array_equal(batch.read_output_rgba8(), one.read_output_rgba8())
    st = batch.stats()
    if aa == 1:  # batched launch: the counters cover its 4 frames (3..6)
        assert st["frames"] == 4 and st["kernel_launches"] == 1
        assert st["samples"] == 4 * W * H
    else:        # aa > 1 is sequential per pixel: frame by frame, stats of the last frame
        assert st["frames"] == 1 and st["active"] == ora.active_list()


def test_unfused_waves_equal_fused_frame_kernel(rv, oracle_mod, cornell):
    """One launch per wave (UNFUSED) and the persistent cooperative k_frame are
    the same computation."""
    from rvpt_b200 import _lib
    eng_u, ora, stats = _render_both(rv, oracle_mod, cornell, 176, 128, CORNELL_POSE, frames=2,
                                     flags=_lib.FLAG_UNFUSED, oracle_flags=0, fov=60.0)
    _assert_bit_equal(eng_u.read_accum_f32(), ora.accum, "unfused")
    assert stats[-1][0]["active"] == stats[-1][1]
    assert stats[-1][0]["kernel_launches"] == 8
    eng_f, _, stats_f = _render_both(rv, oracle_mod, cornell, 176, 128, CORNELL_POSE, frames=2, fov=60.0)
    assert stats_f[-1][0]["kernel_launches"] == 1
    _assert_bit_equal(eng_f.read_accum_f32(), eng_u.read_accum_f32(), "fused vs unfused")


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_tile_partition_reassembles_single_gpu_image(rv, builtin, nranks):
    """Multi-GPU sharding without N GPUs: the ranks' tile sets,
    rendered one after the other on this GPU, add up to the 1-GPU image bit
    for bit — global (x, y, W) seed the RNG (util.glsl:35-36)."""
    W, H = 208, 120  # ragged: 13 x 7.5 tiles
    cam = rv.camera_data(translation=PINNED_POSE, aspect=W / H)
    full = rv.Engine(W, H)
    full.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    acc = np.zeros((H, W, 4), np.float32)
    out = np.zeros((H, W, 4), np.uint8)
    total_samples = 0
    for f in range(2):
        full.render_frame(rv.default_settings(frame=f), cam)
    for r in range(nranks):
        eng = rv.Engine(W, H, rank=r, nranks=nranks)
        eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
        for f in range(2):
            eng.render_frame(rv.default_settings(frame=f), cam)
        a = eng.read_accum_f32()
        o = eng.read_output_rgba8()
        acc += a  # disjoint supports: other ranks' pixels are exactly 0
        out += o
        total_samples += eng.stats()["samples"]
        eng.close()
    _assert_bit_equal(acc, full.read_accum_f32(), f"{nranks}-rank reassembly")
    assert np.array_equal(out, full.read_output_rgba8())
    assert total_samples == W * H


def test_checkpoint_resume_roundtrip(rv, oracle_mod, builtin):
    W, H = 144, 80
    cam = rv.camera_data(translation=PINNED_POSE, aspect=W / H)
    a = rv.Engine(W, H)
    a.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(3):
        a.render_frame(rv.default_settings(frame=f), cam)
    snap = a.read_accum_f32()
    b = rv.Engine(W, H)
    b.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    b.write_accum_f32(snap)
    _assert_bit_equal(b.read_accum_f32(), snap, "write/read accum")
    for eng in (a, b):
        eng.render_frame(rv.default_settings(frame=3), cam)
    _assert_bit_equal(b.read_accum_f32(), a.read_accum_f32(), "resumed frame")


def test_device_arithmetic_matches_host_header(rv, oracle_mod):
    """include/rvpt_math.h evaluated on the device == on the host, bit for bit,
    and -fmad=false was honoured."""
    import ctypes as C
    lib = rv._lib.load()
    ol = oracle_mod.load()
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(0, 2 * np.pi, 20000), rng.uniform(-50, 50, 5000),
                        [0.0, np.pi, 2 * np.pi, 6.2831855]]).astype(np.float32)
    out = np.zeros((len(x), 2), np.float32)
    assert lib.rvpt_b200_selftest_math(0, 0, x.ctypes.data, len(x), out.ctypes.data) == 0
    s = np.zeros(len(x), np.float32)
    c = np.zeros(len(x), np.float32)
    ol.rvpt_oracle_sincos(x.ctypes.data, len(x), s.ctypes.data, c.ctypes.data)
    assert np.array_equal(out[:, 0].view(np.uint32), s.view(np.uint32))
    assert np.array_equal(out[:, 1].view(np.uint32), c.view(np.uint32))

    v = rng.normal(size=(4096, 3)).astype(np.float32)
    nout = np.zeros_like(v)
    nref = np.zeros_like(v)
    assert lib.rvpt_b200_selftest_math(0, 2, v.ctypes.data, len(v), nout.ctypes.data) == 0
    ol.rvpt_oracle_normalize(v.ctypes.data, len(v), nref.ctypes.data)
    assert np.array_equal(nout.view(np.uint32), nref.view(np.uint32))

    probe = np.array([[1.0001220703125, 0.9998779296875, -1.0, np.pi / 2]], np.float32)
    pout = np.zeros((1, 2), np.float32)
    assert lib.rvpt_b200_selftest_math(0, 3, probe.ctypes.data, 1, pout.ctypes.data) == 0
    assert pout[0, 0] == 1.0, "device code was compiled with FMA contraction"

    seeds = rng.integers(1, 2**32, size=1024, dtype=np.uint32)
    rout = np.zeros((len(seeds), 2), np.float32)
    assert lib.rvpt_b200_selftest_math(0, 1, seeds.view(np.float32).ctypes.data, len(seeds),
                                       rout.ctypes.data) == 0
    s0 = seeds.copy()
    for k in range(2):
        s0 ^= s0 << np.uint32(13)
        s0 ^= s0 >> np.uint32(17)
        s0 ^= s0 << np.uint32(5)
        assert np.array_equal(rout[:, k], s0.astype(np.float32) / np.float32(4294967296.0))


def test_unsupported_and_error_paths(rv, builtin):
    eng = rv.Engine(64, 64)
    with pytest.raises(rv.EngineError) as e:
        eng.render_frame(rv.default_settings(), rv.camera_data())
    assert e.value.code == -3  # no scene
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    with pytest.raises(rv.EngineError) as e:
        eng.render_frame(rv.default_settings(aa=0), rv.camera_data())
    assert e.value.code == -1  # the reference divides by aa
    bad = builtin.triangles.copy()
    bad["material_id"][5, 0] = 7
    with pytest.raises(rv.EngineError):
        eng.upload_scene(bad, builtin.materials, builtin.nodes)
    nodes = builtin.nodes.copy()
    nodes["first_child_or_primitive"][0] = 0  # root is its own child: a cycle
    with pytest.raises(rv.EngineError):
        eng.upload_scene(builtin.triangles, builtin.materials, nodes)


def test_full_size_1080p_properties(rv, oracle_mod, builtin):
    """BASELINE config 2 at full size, three progressive frames launched one by one: every
    pixel against the oracle, sample count, and independence of the tile scheduler (two
    engines, same frames -> identical images: no float atomics anywhere)."""
    W, H = 1920, 1080
    cam = rv.camera_data(translation=DEFAULT_POSE, aspect=W / H)
    e1 = rv.Engine(W, H)
    e1.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    e2 = rv.Engine(W, H)
    e2.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(3):
        rs = rv.default_settings(frame=f)
        e1.render_frame(rs, cam)
        e2.render_frame(rs, cam)
        ora.render_frame(rs, cam)
    g = e1.read_accum_f32()
    _assert_bit_equal(g, ora.accum, "1080p, 3 frames, every pixel")
    _assert_bit_equal(g, e2.read_accum_f32(), "determinism across engines")
    st = e1.stats()
    assert st["samples"] == W * H
    assert st["segments"] == sum(st["active"])
    assert st["active"] == ora.active_list()
    assert g[1079].any(), "rows >= 1072 are rendered unless REFERENCE_DISPATCH is set"


@pytest.mark.parametrize("config", ["C2", "C3"])
def test_stated_configs_full_frames_64(rv, oracle_mod, builtin, cornell, config):
    """C2 (built-in scene) and C3 (Cornell box + mesh, mirror and dielectric
    blocks) at 1920x1080, 8 bounces, frames 0..63 — EVERY pixel of the float32 running mean and
    of the rgba8 image against the oracle, through rvpt_b200_render_frames (batched launches)."""
    W, H, N = 1920, 1080, 64
    prep, pose, fov = (builtin, DEFAULT_POSE, 90.0) if config == "C2" else (cornell, CORNELL_POSE, 60.0)
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H)
    eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    eng.render_frames(rv.default_settings(frame=0), cam, N)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
    want_active = np.zeros(64, np.uint64)
    frames_last = eng.stats()["frames"]
    for f in range(N):
        ora.render_frame(rv.default_settings(frame=f), cam)
        if f >= N - frames_last:
            want_active += ora.active
    g = eng.read_accum_f32()
    rel = np.abs(g - ora.accum) / np.maximum(np.abs(ora.accum), 1e-6)
    assert rel.max() <= 1e-4  # BASELINE.json north_star tolerance
    _assert_bit_equal(g, ora.accum, f"{config} after {N} frames, every pixel")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    st = eng.stats()
    got = np.zeros(64, np.uint64)
    got[:len(st["active"])] = st["active"]
    assert np.array_equal(got, want_active), "per-bounce ray counts of the last launch"


def test_cpp_headless_driver_matches_python_host(rv, builtin, tmp_path):
    """f-2: the C++ mirror of main.cpp's loop (load_model -> add_material ->
    initialize -> update/draw) produces the same image as the Python host."""
    import subprocess
    from pathlib import Path
    from rvpt_b200.scene import builtin_mesh, write_obj
    exe = Path(rv.__file__).parent / "rvpt_headless"
    obj, ppm = tmp_path / "bunny.obj", tmp_path / "out.ppm"
    write_obj(obj, *builtin_mesh())
    W, H, frames = 320, 176, 8
    r = subprocess.run([str(exe), str(obj), "--width", str(W), "--height", str(H), "--frames",
                        str(frames), "--translate", "0", "0.8", "-2.5", "--out", str(ppm)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = ppm.read_bytes()
    header = f"P6\n{W} {H}\n255\n".encode()
    assert raw.startswith(header)
    img = np.frombuffer(raw[len(header):], np.uint8).reshape(H, W, 3)

    eng = rv.Engine(W, H)
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    cam = rv.camera_data(translation=PINNED_POSE, aspect=W / H)
    for f in range(frames):  # update(): first frame is 0, then ++ (rvpt.cpp:102-111)
        eng.render_frame(rv.default_settings(frame=f), cam)
    assert np.array_equal(img, eng.read_output_rgba8()[..., :3])


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_other_integrators_bit_exact(rv, oracle_mod, cornell, mode):
    """f-3: binary, color, depth, normal, Utah, AO, Appel, Whitted, Cook
    (integrators.glsl:24-543) against the oracle, 2 frames, aa = 2."""
    eng, ora, _ = _render_both(rv, oracle_mod, cornell, 112, 80, CORNELL_POSE, frames=2, fov=60.0,
                               mode=mode, aa=2, max_bounces=6)
    g, o = eng.read_accum_f32(), ora.accum
    same = (g.view(np.uint32) == o.view(np.uint32)) | (np.isnan(g) & np.isnan(o))
    assert same.all(), f"mode {mode}: {(~same).sum()} words differ"
    assert np.array_equal(eng.read_output_rgba8(), ora.result)


@pytest.mark.parametrize("mode", [10, 11, -1])
def test_hart_sphere_tracer_bit_exact(rv, oracle_mod, cornell, builtin, mode):
    """f-3, eval_integrator's default case (compute_pass.comp:96-97): integrator_Hart
    (integrators.glsl:681-693), the sphere tracer of distance_functions.glsl:36-116 as a heat map
    of its iteration count — every index outside 0..9. Cornell box with aa = 2 over two frames,
    the built-in scene from a pose that sees background, silhouette and surface; then the paths
    that hand the kernel its vertices differently: BVH built inside upload_scene (triangles
    permuted there), the brute-force flag, a second upload on the same engine, a 2-way partition."""
    eng, ora, _ = _render_both(rv, oracle_mod, cornell, 112, 80, CORNELL_POSE, frames=2, fov=60.0, mode=mode, aa=2)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"Hart, mode {mode}, Cornell")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    levels = np.unique(np.rint(ora.accum[..., 0] * 31 * 2))  # mean of 2 x 2 samples of k/31
    assert len(levels) > 8 and ora.accum[..., 0].max() > 1.0  # 32/31: marches that ran out of iterations
    if mode != 10:
        return
    W, H = 96, 64
    cam = rv.camera_data(translation=PINNED_POSE, aspect=W / H)
    rs = rv.default_settings(mode=10)
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
    ora.render_frame(rs, cam)
    # the oracle marches the triangles in buffer order; min() over them is order-independent
    # on these scenes (no NaN distances), so the caller's order and the permuted one agree
    for flags, nodes, tris in ((0, builtin.nodes, builtin.triangles), (0, None, builtin.scene.triangles),
                               (0x4, None, builtin.scene.triangles)):  # 0x4 = RVPT_B200_FLAG_BRUTE_FORCE
        eng = rv.Engine(W, H, flags=flags)
        eng.upload_scene(cornell.triangles, cornell.materials, cornell.nodes)
        eng.render_frame(rs, rv.camera_data(translation=CORNELL_POSE, aspect=W / H, fov=60.0))  # another scene first
        eng.upload_scene(tris, builtin.materials, nodes)
        eng.reset_accum()
        eng.render_frame(rs, cam)
        _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"Hart, built-in scene, flags {flags}, nodes {nodes is not None}")
        eng.close()
    img = np.zeros((H, W, 4), np.float32)
    for r in range(2):
        e = rv.Engine(W, H, rank=r, nranks=2)
        e.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
        e.render_frame(rs, cam)
        img += e.read_accum_f32()  # disjoint supports
        e.close()
    _assert_bit_equal(img, ora.accum, "Hart, 2-way partition")


def test_split_view_with_hart_quadrants(rv, oracle_mod, builtin):
    """Kajiya pixels on the wavefront path, Utah and two Hart quadrants (indices 10 and -3) in
    k_modes, progressive over two frames."""
    W, H = 160, 96
    cam = rv.camera_data(translation=PINNED_POSE, aspect=W / H)
    eng = rv.Engine(W, H)
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(2):
        rs = rv.default_settings(frame=f)
        rs["top_left_render_mode"], rs["top_right_render_mode"] = 10, 9
        rs["bottom_left_render_mode"], rs["bottom_right_render_mode"] = -3, 4
        rs["split_ratio"] = (0.45, 0.5)
        eng.render_frame(rs, cam)
        ora.render_frame(rs, cam)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "split view with Hart")
    assert eng.stats()["active"] == ora.active_list()


def test_split_view_four_integrators(rv, oracle_mod, builtin):
    """compute_pass.comp:134-144: the 4-way split picks an integrator per
    pixel; Kajiya pixels stay on the wavefront path, the others go to k_modes."""
    W, H = 160, 96
    cam = rv.camera_data(translation=PINNED_POSE, aspect=W / H)
    eng = rv.Engine(W, H)
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(3):
        rs = rv.default_settings(frame=f)
        rs["top_left_render_mode"], rs["top_right_render_mode"] = 9, 3
        rs["bottom_left_render_mode"], rs["bottom_right_render_mode"] = 5, 7
        rs["split_ratio"] = (0.3, 0.6)
        eng.render_frame(rs, cam)
        ora.render_frame(rs, cam)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "split view")
    assert eng.stats()["active"] == ora.active_list()  # only Kajiya pixels count as path segments


def test_large_scene_global_memory_path(rv, oracle_mod):
    """f-1: a 20 k-triangle mesh (2.6 MB blob, far above the 192 KB shared-memory
    budget) takes the L2-resident traversal path (kSmem = false) — same results."""
    from conftest import PreparedScene
    prep = PreparedScene(rv, rv.displaced_sphere_scene(20000))
    assert len(prep.nodes) * 32 + len(prep.triangles) * 68 > 192 * 1024
    eng, ora, stats = _render_both(rv, oracle_mod, prep, 192, 128, (0.0, 1.2, -3.0), frames=3, fov=60.0)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "large scene")
    assert stats[-1][0]["active"] == stats[-1][1]
    # the debug integrators use the same path
    eng, ora, _ = _render_both(rv, oracle_mod, prep, 96, 64, (0.0, 1.2, -3.0), fov=60.0, mode=5)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "large scene, AO")
    # a medium scene (~900 triangles, ~150-190 KB with the origin-relative copies) still fits
    # the shared-memory budget of the one-CTA-per-SM kernel
    mid = PreparedScene(rv, rv.displaced_sphere_scene(900))
    eng, ora, _ = _render_both(rv, oracle_mod, mid, 128, 96, (0.0, 1.2, -3.0), frames=2, fov=60.0)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "medium scene in shared memory")


def test_c4_full_size_4k_partition(rv, oracle_mod, builtin):
    """BASELINE config 4 at full size (3840x2160, 8-way tile partition): the
    eight ranks' tile sets, rendered one after the other on this GPU for two
    progressive frames, reassemble the 1-GPU image bit for bit; a band of rows
    is checked against the oracle; sample counts add up."""
    W, H, nranks = 3840, 2160, 8
    cam = rv.camera_data(translation=DEFAULT_POSE, aspect=W / H)
    full = rv.Engine(W, H)
    full.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    full.render_frames(rv.default_settings(frame=0), cam, 2)
    want = full.read_accum_f32()
    want_rgba = full.read_output_rgba8()
    full.close()
    acc = np.zeros((H, W, 4), np.float32)
    rgba = np.zeros((H, W, 4), np.uint8)
    samples = 0
    for r in range(nranks):
        eng = rv.Engine(W, H, rank=r, nranks=nranks)
        eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
        eng.render_frames(rv.default_settings(frame=0), cam, 2)
        acc += eng.read_accum_f32()
        rgba += eng.read_output_rgba8()
        st = eng.stats()
        assert st["frames"] == 2  # one batched launch
        samples += st["samples"]
        eng.close()
    _assert_bit_equal(acc, want, "4K, 8 ranks")
    assert np.array_equal(rgba, want_rgba)
    assert samples == 2 * W * H
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(2):
        ora.render_frame(rv.default_settings(frame=f), cam)
    _assert_bit_equal(want, ora.accum, "4K full frames vs oracle")
    assert np.array_equal(want_rgba, ora.result)


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (16, 16), (17, 33)])
def test_tiny_and_ragged_images(rv, oracle_mod, builtin, size):
    """Edge sizes: a single pixel, less than one tile, exactly one tile, ragged."""
    W, H = size
    eng, ora, stats = _render_both(rv, oracle_mod, builtin, W, H, DEFAULT_POSE, frames=2)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"{W}x{H}")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    assert stats[-1][0]["samples"] == W * H


def test_more_ranks_than_tiles(rv, builtin):
    """A rank that owns no tile renders nothing and reads back zeros."""
    W, H = 24, 16  # 2 tiles
    cam = rv.camera_data(translation=DEFAULT_POSE, aspect=W / H)
    full = rv.Engine(W, H)
    full.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    full.render_frame(rv.default_settings(), cam)
    acc = np.zeros((H, W, 4), np.float32)
    for r in range(4):
        eng = rv.Engine(W, H, rank=r, nranks=4)
        eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
        eng.render_frame(rv.default_settings(), cam)
        a = eng.read_accum_f32()
        if r >= 2:
            assert not a.any() and eng.stats()["samples"] == 0
        acc += a
        eng.close()
    _assert_bit_equal(acc, full.read_accum_f32(), "4 ranks, 2 tiles")


@pytest.mark.parametrize("scene_name", ["builtin", "cornell"])
def test_octant_sorted_nodes_equal_minmax_walk(rv, oracle_mod, builtin, cornell, scene_name):
    """The per-octant (near, far) node copies are an instruction-count optimisation of
    intersect_aabb (intersection.glsl:327-357): same image with and without them, and both
    equal to the oracle. The cornell pose looks straight down +z at axis-aligned walls
    (zero-thickness boxes); the ortho camera's direction has exact zeros (slow path)."""
    from rvpt_b200 import _lib
    prep = builtin if scene_name == "builtin" else cornell
    pose = PINNED_POSE if scene_name == "builtin" else CORNELL_POSE
    for cam_mode in (0, 1):
        a, ora, st_a = _render_both(rv, oracle_mod, prep, 208, 160, pose, frames=3, fov=60.0,
                                    camera_mode=cam_mode)
        b, _, st_b = _render_both(rv, oracle_mod, prep, 208, 160, pose, frames=3, fov=60.0,
                                  flags=_lib.FLAG_NO_OCTANTS, oracle_flags=0, camera_mode=cam_mode)
        _assert_bit_equal(a.read_accum_f32(), ora.accum, "octant copies vs oracle")
        _assert_bit_equal(b.read_accum_f32(), ora.accum, "min/max walk vs oracle")
        assert st_a[-1][0]["active"] == st_b[-1][0]["active"] == st_a[-1][1]


def test_frame_kernel_timeline(rv, builtin):
    """set_timeline: per-CTA phase stamps are monotone and cover the frame."""
    W, H = 640, 360
    cam = rv.camera_data(translation=DEFAULT_POSE, aspect=W / H)
    eng = rv.Engine(W, H)
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    eng.set_timeline(True)
    eng.render_frame(rv.default_settings(frame=0), cam)
    tl = eng.timeline()
    assert tl.shape[0] >= 1 and tl.shape[1] == 16
    assert (tl[:, 0] > 0).all() and (tl[:, 1] >= tl[:, 0]).all() and (tl[:, 2] >= tl[:, 1]).all()
    span_us = (tl.max() - tl[:, 0].min()) / 1e3
    assert 0 < span_us < 1e5
    eng.set_timeline(False)
    eng.render_frame(rv.default_settings(frame=1), cam)
    assert eng.timeline().size == 0


def test_wave_forecast_is_scheduling_only(rv, oracle_mod, builtin):
    """k_frame lets the survivors of a wave run on in their threads when the previous frame
    says the next wave would be tiny. The pinned pose has such waves ([.., 3963, 345, 64, ..]
    at 256x256); moving the camera between frames makes the forecast wrong on purpose. Images
    and per-bounce counts equal the oracle's with and without the forecast."""
    from rvpt_b200 import _lib
    W = H = 256
    poses = [PINNED_POSE, PINNED_POSE, PINNED_POSE, DEFAULT_POSE, DEFAULT_POSE, PINNED_POSE]
    for flags in (0, _lib.FLAG_NO_FORECAST):
        eng = rv.Engine(W, H, flags=flags)
        eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
        ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
        frame = 0
        for i, pose in enumerate(poses):
            if i and pose != poses[i - 1]:
                frame = 0  # rvpt.cpp:102-107: a camera change restarts the accumulation
            cam = rv.camera_data(translation=pose, aspect=W / H)
            rs = rv.default_settings(frame=frame)
            eng.render_frame(rs, cam)
            ora.render_frame(rs, cam)
            assert eng.stats()["active"] == ora.active_list(), (flags, i)
            frame += 1
        _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"forecast flags={flags}")


@pytest.mark.parametrize("scene_name", ["builtin", "cornell", "cornell_on_floor"])
def test_front_to_back_order_equals_reference_order(rv, oracle_mod, builtin, cornell, scene_name):
    """Scenes without coincident faces are walked front to back per direction octant — primary
    rays always, bounce rays from the second frame on when most of them hit something (the
    closed box) — while RVPT_B200_FLAG_REFERENCE_ORDER and scenes WITH coincident faces (Cornell
    blocks standing on the floor with their bottom faces in its plane) use the reference's
    child order (intersection.glsl:402-406). Every variant must equal the oracle, which walks the
    reference's order: the nearest accepted hit is order-independent when no two faces coincide."""
    from conftest import PreparedScene
    from rvpt_b200 import _lib
    if scene_name == "builtin":
        prep, pose, fov, want_order = builtin, PINNED_POSE, 90.0, 1
    elif scene_name == "cornell":
        prep, pose, fov, want_order = cornell, CORNELL_POSE, 60.0, 1
    else:
        prep = PreparedScene(rv, rv.cornell_scene(block_gap=0.0))  # block bottoms in the floor plane
        pose, fov, want_order = CORNELL_POSE, 60.0, 0
    a, ora, st_a = _render_both(rv, oracle_mod, prep, 320, 192, pose, frames=4, fov=fov)
    b, _, st_b = _render_both(rv, oracle_mod, prep, 320, 192, pose, frames=4, fov=fov,
                              flags=_lib.FLAG_REFERENCE_ORDER, oracle_flags=0)
    assert st_a[-1][0]["traversal_order"] == want_order and st_b[-1][0]["traversal_order"] == 0
    _assert_bit_equal(a.read_accum_f32(), ora.accum, "default order vs oracle")
    _assert_bit_equal(b.read_accum_f32(), ora.accum, "reference order vs oracle")
    for f in range(4):
        assert st_a[f][0]["active"] == st_b[f][0]["active"] == st_a[f][1]


@pytest.mark.parametrize("unfused", [False, True])
def test_octant_sorted_queues_are_scheduling_only(rv, oracle_mod, cornell, unfused):
    """In closed scenes every wave's rays are sorted into bins (direction octant x origin cell),
    so the rays a warp loads together walk the same node array along similar paths. Which warp
    traces a path never changes the path: images and per-bounce counts equal the oracle's with
    and without the sorting (one-launch-per-wave never sorts), incl. a spread tail (small image)
    and aa passes."""
    from rvpt_b200 import _lib
    base = _lib.FLAG_UNFUSED if unfused else 0
    for W, H, kw in ((200, 152, dict(frames=3)), (48, 40, dict(frames=2, aa=2)), (640, 360, dict(frames=2))):
        a, ora, st_a = _render_both(rv, oracle_mod, cornell, W, H, CORNELL_POSE, fov=60.0, flags=base,
                                    oracle_flags=0, **kw)
        b, _, st_b = _render_both(rv, oracle_mod, cornell, W, H, CORNELL_POSE, fov=60.0,
                                  flags=base | _lib.FLAG_NO_QUEUE_SORT, oracle_flags=0, **kw)
        _assert_bit_equal(a.read_accum_f32(), ora.accum, f"sorted queues {W}x{H}")
        _assert_bit_equal(b.read_accum_f32(), ora.accum, f"single queue {W}x{H}")
        assert st_a[-1][0]["active"] == st_b[-1][0]["active"] == st_a[-1][1]


# ---- batched launches: rvpt_b200_render_frames merges the waves of consecutive frames --------

def _batched_vs_oracle(rv, oracle_mod, prep, W, H, pose, batches, flags=0, fov=90.0, rank=0, nranks=1,
                       **settings_kw):
    """Renders `batches` (a list of frame counts) with consecutive render_frames calls and the
    same frames one by one with the oracle; returns (engine, oracle, per-launch oracle ray sums)."""
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H, flags=flags, rank=rank, nranks=nranks)
    eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes, flags=flags & 0x3)
    frame = 0
    for n in batches:
        eng.render_frames(rv.default_settings(frame=frame, **settings_kw), cam, n)
        st = eng.stats()
        want = np.zeros(64, np.uint64)
        for f in range(frame, frame + n):
            ora.render_frame(rv.default_settings(frame=f, **settings_kw), cam)
            if f >= frame + n - st["frames"]:
                want += ora.active
        if nranks == 1:
            got = np.zeros(64, np.uint64)
            got[:len(st["active"])] = st["active"]
            assert np.array_equal(got, want), f"ray counts of the last launch ({st['frames']} frames)"
            assert st["samples"] == int(want[0]) or settings_kw.get("max_bounces", 8) == 0
        frame += n
    return eng, ora


@pytest.mark.parametrize("case", ["builtin_16", "pinned_5_11", "cornell_8_8", "cornell_rgba8", "dispatch",
                                  "ortho", "spherical", "b1", "b2", "b16", "max_batch", "two_launches"])
def test_batched_launch_bit_exact(rv, oracle_mod, builtin, cornell, case):
    """A batch renders frames f..f+n-1 in ONE launch: primary rays of all frames in one wave,
    bounce waves merged across frames, samples folded into the running mean in frame order by
    the resolve phase. The images after the batch equal n render_frame calls / the oracle."""
    from rvpt_b200 import _lib
    kw = {}
    if case == "builtin_16":
        prep, W, H, pose, batches = builtin, 320, 180, DEFAULT_POSE, [16]
    elif case == "pinned_5_11":
        prep, W, H, pose, batches = builtin, 256, 256, PINNED_POSE, [5, 11, 1, 2]
    elif case == "cornell_8_8":   # the second launch queues by octant (closed-scene forecast)
        prep, W, H, pose, batches, kw = cornell, 200, 152, CORNELL_POSE, [8, 8], dict(fov=60.0)
    elif case == "cornell_rgba8":
        prep, W, H, pose, batches, kw = cornell, 160, 96, CORNELL_POSE, [6, 3], dict(
            fov=60.0, flags=_lib.FLAG_ACCUM_RGBA8)
    elif case == "dispatch":
        prep, W, H, pose, batches, kw = builtin, 200, 90, PINNED_POSE, [4], dict(
            flags=_lib.FLAG_REFERENCE_DISPATCH)
    elif case == "ortho":
        prep, W, H, pose, batches, kw = builtin, 128, 64, PINNED_POSE, [3], dict(camera_mode=1)
    elif case == "spherical":
        prep, W, H, pose, batches, kw = builtin, 128, 64, PINNED_POSE, [3], dict(camera_mode=2)
    elif case in ("b1", "b2", "b16"):
        prep, W, H, pose, batches, kw = cornell, 64, 48, CORNELL_POSE, [3, 2], dict(
            fov=60.0, max_bounces=int(case[1:]))
    elif case == "max_batch":
        prep, W, H, pose, batches = builtin, 17, 33, DEFAULT_POSE, [64]
    else:  # more frames than one launch can tag (64): split into equal launches
        prep, W, H, pose, batches = builtin, 64, 48, PINNED_POSE, [70]
    eng, ora = _batched_vs_oracle(rv, oracle_mod, prep, W, H, pose, batches, **kw)
    if kw.get("flags", 0) & _lib.FLAG_ACCUM_RGBA8:
        assert np.array_equal(eng.read_accum_f32(), ora.accum_f32())
    else:
        _assert_bit_equal(eng.read_accum_f32(), ora.accum, case)
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    if case == "two_launches":
        assert eng.stats()["frames"] == 35


def test_leaf_lists_of_the_batched_primary_wave(rv, oracle_mod, builtin, cornell):
    """Batched launches render their primary wave pixel block by pixel block: per 8x4 block the
    leaves its beam can enter are listed once and the block's rays of a whole group of frames
    test those leaf boxes instead of walking the tree (kernels.cu, primary_phase_beam). The list
    must be complete and the result the walk's, bit for bit: against the oracle and against the
    same engine with RVPT_B200_FLAG_NO_LEAF_LISTS, for poses that exercise every branch —
    camera inside the scene box, far away (a block sees more than 32 leaves: ordinary walk),
    rotated (blocks whose direction components change sign: ordinary walk), grazing along a wall,
    wide and narrow fields of view, ragged images, frame groups that do not divide the batch,
    a tile partition — and for a BVH whose boxes are NOT nested (a child box grown beyond its
    parent's: legal in the reference's format, lists must switch themselves off)."""
    from rvpt_b200 import _lib
    cases = [
        (builtin, 320, 180, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 90.0, 19),
        (builtin, 256, 144, (0.0, 0.8, -2.5), (0.0, 0.0, 0.0), 90.0, 33),
        (builtin, 200, 120, (0.0, 1.0, -14.0), (0.0, 0.0, 0.0), 30.0, 17),    # tiny bunny: > 32 leaves per block
        (builtin, 208, 120, (0.6, 0.9, -1.2), (0.3, -0.5, 0.2), 120.0, 16),    # rotated, wide
        (builtin, 97, 61, (-0.4, 0.4, 0.9), (0.1, 2.6, 0.0), 70.0, 7),
        (cornell, 240, 136, (0.0, 1.2, -3.4), (0.0, 0.0, 0.0), 60.0, 18),
        (cornell, 160, 96, (0.95, 0.05, -0.9), (0.02, -0.1, 0.0), 100.0, 16),   # grazing along wall and floor
        (cornell, 160, 96, (0.0, 1.0, 0.0), (1.3, 0.4, 0.0), 150.0, 5),        # inside the box
    ]
    for prep, W, H, pose, rot, fov, n in cases:
        cam = rv.camera_data(translation=pose, rotation=rot, aspect=W / H, fov=fov)
        ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
        for f in range(n):
            ora.render_frame(rv.default_settings(frame=f), cam)
        for flags in (0, _lib.FLAG_NO_LEAF_LISTS):
            eng = rv.Engine(W, H, flags=flags)
            eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
            eng.render_frames(rv.default_settings(frame=0), cam, n)
            assert eng.stats()["kernel_launches"] == 1
            _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"pose {pose} rot {rot} fov {fov} flags {flags:#x}")
            eng.close()
    # tile partition: the blocks of a rank are the same blocks
    prep, W, H, pose, rot, fov, n = cases[3]
    cam = rv.camera_data(translation=pose, rotation=rot, aspect=W / H, fov=fov)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
    for f in range(n):
        ora.render_frame(rv.default_settings(frame=f), cam)
    img = np.zeros((H, W, 4), np.float32)
    for r in range(3):
        e = rv.Engine(W, H, rank=r, nranks=3)
        e.upload_scene(prep.triangles, prep.materials, prep.nodes)
        e.render_frames(rv.default_settings(frame=0), cam, n)
        img += e.read_accum_f32()
        e.close()
    _assert_bit_equal(img, ora.accum, "leaf lists, 3-way partition")
    # boxes that are not nested: grow one leaf's box far beyond its parent's
    nodes = builtin.nodes.copy()
    leaf = int(np.flatnonzero(nodes["primitive_count"] > 0)[7])
    nodes["bounds"][leaf] += np.array([-3, 3, -3, 3, -3, 3], np.float32)
    W, H, n = 192, 112, 9
    cam = rv.camera_data(translation=(0.0, 0.8, -2.5), aspect=W / H)
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, nodes)
    for f in range(n):
        ora.render_frame(rv.default_settings(frame=f), cam)
    eng = rv.Engine(W, H)
    eng.upload_scene(builtin.triangles, builtin.materials, nodes)
    eng.render_frames(rv.default_settings(frame=0), cam, n)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "boxes not nested")


def test_batched_launch_partition_and_large_scene(rv, oracle_mod, builtin):
    """Batches on a 3-way tile partition (ragged image) reassemble the oracle's image; a scene on
    the global-memory path batches too; NO_BATCH is the same computation frame by frame."""
    from conftest import PreparedScene
    from rvpt_b200 import _lib
    W, H = 208, 120
    acc = np.zeros((H, W, 4), np.float32)
    for r in range(3):
        eng, ora = _batched_vs_oracle(rv, oracle_mod, builtin, W, H, PINNED_POSE, [7], rank=r, nranks=3)
        acc += eng.read_accum_f32()
        eng.close()
    _assert_bit_equal(acc, ora.accum, "3-way partition, batch of 7")
    prep = PreparedScene(rv, rv.displaced_sphere_scene(20000))
    eng, ora = _batched_vs_oracle(rv, oracle_mod, prep, 192, 128, (0.0, 1.2, -3.0), [4, 2], fov=60.0)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "global-memory path, batches of 4 and 2")
    eng, ora = _batched_vs_oracle(rv, oracle_mod, builtin, 160, 96, DEFAULT_POSE, [5],
                                  flags=_lib.FLAG_NO_BATCH)
    assert eng.stats()["frames"] == 1
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "NO_BATCH")


def test_batch_size_follows_the_queue_budget(rv, oracle_mod, builtin, monkeypatch):
    """The path queues of a launch must fit RVPT_B200_QUEUE_BUDGET_MIB: 256x256 needs 9.75 MiB per
    frame of a batch (2 queues x 64 B + 12 B sort key / permutation + 16 B parked sample per
    pixel), so 35 MiB allows 3 frames."""
    monkeypatch.setenv("RVPT_B200_QUEUE_BUDGET_MIB", "35")
    eng, ora = _batched_vs_oracle(rv, oracle_mod, builtin, 256, 256, DEFAULT_POSE, [7])
    assert eng.stats()["frames"] == 1  # 7 frames -> launches of 3, 3, 1
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "budget-limited batches")


@pytest.mark.parametrize("flags_name", ["batched", "frame_by_frame"])
def test_cuda_graph_replay_is_safe_for_any_launch_count(rv, oracle_mod, builtin, flags_name):
    """A CUDA graph captured through the C ABI replays correctly whatever number of launches it
    holds (3 frame launches here, or one batched launch): no device state depends on host-side
    launch parity — the last CTA of every launch leaves the counters clean (round-1 advice)."""
    torch = pytest.importorskip("torch")
    from rvpt_b200 import _lib
    W, H, N = 320, 180, 3
    cam = rv.camera_data(translation=DEFAULT_POSE, aspect=W / H)
    rs = rv.default_settings(frame=0)
    flags = 0 if flags_name == "batched" else _lib.FLAG_NO_BATCH
    eng = rv.Engine(W, H, flags=flags)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    eng.render_frames(rs, cam, N)  # warm-up: allocations happen outside the capture
    eng.sync()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            eng.render_frames(rs, cam, N)
    ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
    for f in range(N):
        ora.render_frame(rv.default_settings(frame=f), cam)
    for replay in range(3):
        g.replay()
        torch.cuda.synchronize()
        _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"graph replay {replay}")
    eng.set_stream(None)
    eng.close()


@pytest.mark.parametrize("case", ["pinned", "cornell"])
def test_full_size_repeated_launches_stay_exact(rv, oracle_mod, builtin, cornell, case):
    """Full-size batches launched repeatedly on one engine, so that every scheduling mode the
    previous launch's statistics switch on is in effect (wave forecast, front-to-back bounce
    arrays, binned ray sort): every pixel equals the oracle after each repetition. At this scale
    (1e8 rays) some rays do hit a shared edge or a corner within rounding distance — where the
    front-to-back walk alone would pick another triangle than the reference's walk (round 2 found
    2 such pixels in the pinned pose, 1 in the Cornell box); the ambiguity re-trace must catch them."""
    W, H = 1920, 1080
    prep, pose, fov, n = (builtin, PINNED_POSE, 90.0, 64) if case == "pinned" else (cornell, CORNELL_POSE, 60.0, 16)
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
    for f in range(n):
        ora.render_frame(rv.default_settings(frame=f), cam)
    eng = rv.Engine(W, H)
    eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    for rep in range(3):
        eng.render_frames(rv.default_settings(frame=0), cam, n)
        _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"{case}, repetition {rep}")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)


def test_tridel_interior_560k_triangles(rv, oracle_mod):
    """f-1: the reference's large model (assets/models/tridel-interior-test.obj, 560 021
    triangles, packed by tools/pack_tridel.py) through the library's own BVH builder and the
    L2-resident traversal path: 640x360, frames 0..3 in one batched launch, every pixel against
    the oracle; the per-bounce ray counts too."""
    from rvpt_b200.scene import TRIDEL_NPZ
    if not TRIDEL_NPZ.exists():
        pytest.skip("rvpt_b200/assets/tridel_interior.npz not packed (tools/pack_tridel.py needs the reference tree)")
    from conftest import PreparedScene
    scene, pose, fov = rv.tridel_scene()
    prep = PreparedScene(rv, scene)
    assert len(prep.triangles) == 560021
    W, H, N = 640, 360, 4
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H)
    eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    eng.render_frames(rv.default_settings(frame=0), cam, N)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
    want = np.zeros(64, np.uint64)
    for f in range(N):
        ora.render_frame(rv.default_settings(frame=f), cam)
        want += ora.active
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, "tridel interior, 4 frames")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    st = eng.stats()
    got = np.zeros(64, np.uint64)
    got[:len(st["active"])] = st["active"]
    assert np.array_equal(got, want)
    assert len(st["active"]) == 8 and st["active"][7] > 0, "an interior: paths live through all 8 bounces"


def test_async_double_buffered_readback(rv, oracle_mod, builtin):
    """rvpt_b200_read_output_rgba8_async: the copy of one step's image overlaps the frames of the
    next, which fill the second raster image — both copies arrive intact, in any interleaving with
    the synchronous read."""
    torch = pytest.importorskip("torch")
    W, H = 640, 360
    cam = rv.camera_data(translation=DEFAULT_POSE, aspect=W / H)
    want = {}
    for n in (3, 8):
        ora = oracle_mod.OracleRenderer(W, H, builtin.triangles, builtin.materials, builtin.nodes)
        for f in range(n):
            ora.render_frame(rv.default_settings(frame=f), cam)
        want[n] = ora.result.copy()
    eng = rv.Engine(W, H)
    eng.upload_scene(builtin.triangles, builtin.materials, builtin.nodes)
    bufs = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    for rep in range(3):
        eng.render_frames(rv.default_settings(frame=0), cam, 3)
        eng.wait_output()
        eng.read_output_rgba8_async(bufs[0])
        eng.render_frames(rv.default_settings(frame=0), cam, 8)      # fills the other image meanwhile
        assert np.array_equal(eng.read_output_rgba8(), want[8])     # synchronous read: the latest frames
        eng.read_output_rgba8_async(bufs[1])
        eng.wait_output()
        assert np.array_equal(bufs[0], want[3]), f"first copy, repetition {rep}"
        assert np.array_equal(bufs[1], want[8]), f"second copy, repetition {rep}"
    eng.close()


@pytest.mark.parametrize("scene_name", ["builtin", "cornell", "mesh20k", "identical", "tridel"])
def test_gpu_bvh_builder(rv, oracle_mod, scene_name):
    """f-1: the GPU builder (linear BVH: Morton codes, radix sort, Karras hierarchy, bottom-up fit;
    rvpt_b200/csrc/bvh_gpu.cu) emits a valid tree in the reference's node format — every triangle in
    exactly one leaf, boxes nested, children adjacent — and rendering with it equals the oracle
    walking the same nodes (and the image of the SAH tree: the nearest hit does not depend on the BVH)."""
    from conftest import PreparedScene
    from test_abi_and_host import _check_bvh
    from rvpt_b200 import _lib
    from rvpt_b200.scene import TRIDEL_NPZ, make_triangles
    pose, fov = PINNED_POSE, 90.0
    if scene_name == "builtin":
        scene = rv.builtin_scene()
    elif scene_name == "cornell":
        scene, pose, fov = rv.cornell_scene(), CORNELL_POSE, 60.0
    elif scene_name == "mesh20k":
        scene, pose, fov = rv.displaced_sphere_scene(20000), (0.0, 1.2, -3.0), 60.0
    elif scene_name == "identical":  # equal Morton codes everywhere: the index bits split
        base = np.random.default_rng(3).normal(size=(1, 3, 3)).astype(np.float32)
        v = np.repeat(base, 37, axis=0)
        scene = rv.Scene(make_triangles(v[:, 0], v[:, 1], v[:, 2], 0), rv.make_material((0.7, 0.7, 0.7, 0)))
        pose = (0.0, 0.0, -4.0)
    else:
        if not TRIDEL_NPZ.exists():
            pytest.skip("rvpt_b200/assets/tridel_interior.npz not packed")
        scene, pose, fov = rv.tridel_scene()
    nodes, perm, ms = rv.build_bvh_gpu(scene.triangles)
    n = len(scene.triangles)
    assert len(nodes) == 2 * n - 1 and ms > 0.0
    if n <= 20000:
        import sys
        sys.setrecursionlimit(100000)
        leaves, depth = _check_bvh(nodes, perm, scene.triangles)
        assert leaves == n and depth < 63
    else:
        assert sorted(perm.tolist()) == list(range(n))
        assert int((nodes["primitive_count"] > 0).sum()) == n
    tris = np.ascontiguousarray(scene.triangles[perm])
    W, H = 256, 144
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    eng = rv.Engine(W, H)
    eng.upload_scene(tris, scene.materials, nodes)
    ora = oracle_mod.OracleRenderer(W, H, tris, scene.materials, nodes)
    for f in range(2):
        rs = rv.default_settings(frame=f)
        eng.render_frame(rs, cam)
        ora.render_frame(rs, cam)
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"GPU-built BVH, {scene_name}")
    assert eng.stats()["active"] == ora.active_list()
    # upload_scene(nodes=NULL) with RVPT_B200_FLAG_GPU_BVH builds the same tree inside the library
    eng2 = rv.Engine(W, H, flags=_lib.FLAG_GPU_BVH)
    eng2.upload_scene(scene.triangles, scene.materials, None)
    for f in range(2):
        eng2.render_frame(rv.default_settings(frame=f), cam)
    _assert_bit_equal(eng2.read_accum_f32(), ora.accum, f"GPU BVH built inside upload_scene, {scene_name}")
    print(f"GPU BVH build, {scene_name}: {n} triangles in {ms:.2f} ms (device)")


@pytest.mark.parametrize("seed", range(40))
def test_randomised_configurations(rv, oracle_mod, builtin, cornell, seed):
    """Seeded random walks through the configuration space — image size (ragged), scene, pose,
    camera, bounce limit, aa, accumulation mode, dispatch rule, implementation flags, tile
    partition, and a mix of render_frame / render_frames calls of random lengths on ONE engine —
    each compared with the oracle fed the same frame sequence."""
    from rvpt_b200 import _lib
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(1, 200)), int(rng.integers(1, 120))
    prep, pose, fov = (builtin, [DEFAULT_POSE, PINNED_POSE][int(rng.integers(2))], 90.0) if rng.random() < 0.5 \
        else (cornell, CORNELL_POSE, 60.0)
    kw = dict(max_bounces=int(rng.choice([1, 2, 3, 8, 8, 8, 16])), aa=int(rng.choice([1, 1, 1, 2, 3])),
              camera_mode=int(rng.choice([0, 0, 0, 1, 2])))
    split = None  # now and then another integrator everywhere, or the 4-way split view
    if rng.random() < 0.15:
        kw["mode"] = int(rng.integers(0, 9))
    elif rng.random() < 0.15:
        split = ([int(v) for v in rng.choice([9, 9, 0, 1, 3, 4, 5, 6, 7, 8], 4)], (float(rng.random()), float(rng.random())))

    def settings(f):
        rs = rv.default_settings(frame=f, **kw)
        if split:
            (rs["top_left_render_mode"], rs["top_right_render_mode"], rs["bottom_left_render_mode"],
             rs["bottom_right_render_mode"]) = split[0]
            rs["split_ratio"] = split[1]
        return rs
    accum_flags = int(rng.choice([0, 0, _lib.FLAG_ACCUM_RGBA8])) | int(rng.choice([0, 0, _lib.FLAG_REFERENCE_DISPATCH]))
    impl_flags = 0
    for f in (_lib.FLAG_NO_OCTANTS, _lib.FLAG_NO_FORECAST, _lib.FLAG_REFERENCE_ORDER, _lib.FLAG_NO_QUEUE_SORT,
              _lib.FLAG_NO_BATCH, _lib.FLAG_UNFUSED):
        if rng.random() < 0.2:
            impl_flags |= f
    nranks = int(rng.choice([1, 1, 2, 3]))
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes, flags=accum_flags)
    engines = []
    for r in range(nranks):
        e = rv.Engine(W, H, flags=accum_flags | impl_flags, rank=r, nranks=nranks)
        e.upload_scene(prep.triangles, prep.materials, prep.nodes)
        engines.append(e)
    frame = 0
    for _ in range(int(rng.integers(1, 5))):
        n = int(rng.choice([1, 1, 2, 5, 9]))
        batched = rng.random() < 0.6
        for e in engines:
            if batched:
                e.render_frames(settings(frame), cam, n)
            else:
                for f in range(frame, frame + n):
                    e.render_frame(settings(f), cam)
        for f in range(frame, frame + n):
            ora.render_frame(settings(f), cam)
        frame += n
    what = f"seed {seed}: {W}x{H} {kw} split {split} accum {accum_flags:#x} impl {impl_flags:#x} ranks {nranks}"
    accs = [e.read_accum_f32() for e in engines]
    acc = accs[0].copy()
    for a in accs[1:]:                                   # disjoint supports (NaNs of the depth view stay put)
        acc = np.where(a.view(np.uint32) != 0, a, acc)
    rgba = sum(e.read_output_rgba8().astype(np.uint16) for e in engines).astype(np.uint8)
    if accum_flags & _lib.FLAG_ACCUM_RGBA8:
        assert np.array_equal(acc, ora.accum_f32()), what
    else:
        want = ora.accum
        same = (acc.view(np.uint32) == want.view(np.uint32)) | (np.isnan(acc) & np.isnan(want))
        assert same.all(), f"{what}: {(~same).sum()} words differ"
    assert np.array_equal(rgba, ora.result), what
    for e in engines:
        e.close()


@pytest.mark.parametrize("seed", range(24))
def test_random_camera_poses(rv, oracle_mod, builtin, cornell, seed):
    """The front-to-back walk is exact for every view, not just the three poses of the stated
    configurations: seeded random camera positions (inside, outside and on the walls of the scene's
    box), orientations (all three angles) and fields of view; 640x360 x 6 frames (1920x1080 x 16
    for the last four) in two batched launches (the second one runs with the forecast, the ordered
    bounce arrays and the octant queues), compared with the oracle bit for bit — running mean,
    result image, ray counts."""
    rng = np.random.default_rng(7000 + seed)
    prep = cornell if seed % 2 else builtin
    v = np.concatenate([prep.triangles[k][:, :3] for k in ("vertex0", "vertex1", "vertex2")])
    lo, hi = v.min(0), v.max(0)
    kind = seed % 3
    big = seed >= 20                                               # four of them at the stated size
    (W, H), batches = ((1920, 1080), (4, 12)) if big else ((640, 360), (2, 4))
    for attempt in range(40):                                      # until the view shows the scene
        pos = lo + (hi - lo) * rng.random(3)                       # inside the bounding box
        if kind == 1:
            pos = lo - 0.5 * (hi - lo) + 2.0 * (hi - lo) * rng.random(3)  # anywhere around it
        elif kind == 2:
            pos[int(rng.integers(3))] = [lo, hi][int(rng.integers(2))][int(rng.integers(3))]  # on a bounding plane
        rot = rng.uniform(-180.0, 180.0, 3) * np.array([1.0, 0.5, 0.25])
        cam = rv.camera_data(translation=pos.astype(np.float32), rotation=rot.astype(np.float32), aspect=W / H,
                             fov=float(rng.uniform(25.0, 120.0)))
        probe = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
        probe.render_frame(rv.default_settings(frame=0), cam)
        if probe.active[1] * 10 > W * H:
            break
    else:
        pytest.fail("no pose found that sees the scene")
    eng = rv.Engine(W, H)
    eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    ora = oracle_mod.OracleRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
    frame = 0
    for n in batches:
        eng.render_frames(rv.default_settings(frame=frame), cam, n)
        want = np.zeros(64, np.uint64)
        for f in range(frame, frame + n):
            ora.render_frame(rv.default_settings(frame=f), cam)
            want += ora.active
        frame += n
        got = np.zeros(64, np.uint64)
        st = eng.stats()
        got[:len(st["active"])] = st["active"]
        assert st["frames"] == n and np.array_equal(got, want), f"seed {seed}: ray counts"
    _assert_bit_equal(eng.read_accum_f32(), ora.accum, f"seed {seed}: pose {pos} rotation {rot}")
    assert np.array_equal(eng.read_output_rgba8(), ora.result)
    eng.close()


@pytest.mark.parametrize("config", ["C2", "C3"])
def test_gpu_against_the_reference_shader_compiled_for_the_host(rv, builtin, cornell, config):
    """The CUDA path against the reference's OWN implementation at full size, live: the shipped
    compute_pass.comp.spv translated to C++ (oracle/spirv_to_cpp.py) and compiled for the host
    (oracle/_ref/libref_shader.so — built next to the reference tree, it travels to this box).
    1920x1080, four progressive frames in one batched launch, every pixel, float32 bit for bit."""
    from oracle import ref_shader
    if not ref_shader.LIB_PATH.exists():
        pytest.skip("oracle/_ref/libref_shader.so was not built (needs the reference tree)")
    W, H, N = 1920, 1080, 4
    prep, pose, fov = (builtin, DEFAULT_POSE, 90.0) if config == "C2" else (cornell, CORNELL_POSE, 60.0)
    cam = rv.camera_data(translation=pose, aspect=W / H, fov=fov)
    ref = ref_shader.RefShaderRenderer(W, H, prep.triangles, prep.materials, prep.nodes)
    for f in range(N):
        ref.render_frame(rv.default_settings(frame=f), cam)
    eng = rv.Engine(W, H)
    eng.upload_scene(prep.triangles, prep.materials, prep.nodes)
    for rep in range(2):  # the second launch runs with the forecast / ordered bounce arrays / octant queues on
        eng.render_frames(rv.default_settings(frame=0), cam, N)
        got = eng.read_accum_f32()
        same = got[..., :3].view(np.uint32) == ref.temporal[..., :3].view(np.uint32)
        assert same.all(), f"{config}, launch {rep}: {(~same).sum()} float32 words differ from the reference's shader"
        assert np.array_equal(eng.read_output_rgba8()[..., :3], ref.result_rgba8()[..., :3])
