#!/usr/bin/env python
"""bench.py — Msamples/s of the path-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): built-in scene (143-triangle bunny,
white Lambert, procedural sky; src/rvpt/main.cpp:102-107), literal default camera, 1920x1080,
max_bounces 8, aa 1, progressive accumulation. One STEP is the progressive batch of 8(d):
frames 0..63 (`--frames`), through ONE C-ABI call, rvpt_b200_render_frames — every step restarts
the running mean (compute_pass.comp:146-148 multiplies the previous image by min(frame, 1)).
The workload's bytes (BVH, triangles, materials, camera) are committed under oracle/workloads/.

  value   device-resident throughput: scene uploaded once, CUDA events on the launch stream
          around each step, 256 MiB memset between steps (L2 flush, outside the events).
  e2e     the same step through the host-buffer API: upload_scene from host arrays + the
          render_frames call (settings + camera by value) + the rgba8 image read back into pinned
          host memory (double-buffered: the copy of step k overlaps the frames of step k+1) — all
          inside the timed region, wall clock.
  parity_ok  the rgba8 image the last timed step left behind == the CPU oracle's image of the
          same 64 frames, byte for byte (at every N: the image rank 0 assembled).
  c4, c3  secondary records: BASELINE.json configs[3] (3840x2160 x 16 progressive frames) and
          configs[2] (Cornell box + mesh, mirror / dielectric blocks, 1920x1080 x 16 frames), same
          timing discipline and the same parity check.

N > 1 (one process per GPU): the frame is sharded by 16x16 pixel tile (tile_id % N == rank);
every rank renders its tiles, nothing is exchanged while rendering. Assembly on rank 0:
  p2p (default)  the kernels' resolve phase stores finished rgba8 pixels straight into rank 0's
                 raster image over NVLink peer memory (CUDA IPC) — fused, no collective; NCCL
                 carries the rendezvous and one barrier per step;
  nccl           the kernels write tiles into the rank's slot of a gather buffer, ONE
                 all_gather_into_tensor per step, untile on rank 0.
Total work is fixed ("scaling": "strong").

--impl reference times the CPU restatement of the reference shader (oracle/rvpt_oracle.cpp; the
reference's Vulkan path cannot run here) with all host threads, one full frame of the
same workload per step. That arm loads only oracle/ — never the product library.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Msamples/s at 1920x1080, 8-bounce, built-in scene"
UNIT = "Msamples/s"


def measured_hbm_peak():
    """HBM roofline denominator: the driver-written MEASURED_PEAKS.json when present, else the
    fallback of /opt/skills/guides/B200_PROFILING.md."""
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        if float(peaks.get("hbm_gbs", 0.0)) > 0.0:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=64, help="progressive frames per step (SURVEY 8(d): 64)")
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--aa", type=int, default=1)
    ap.add_argument("--scene", default="builtin", choices=["builtin", "cornell", "mesh", "tridel"])
    ap.add_argument("--mesh-tris", type=int, default=500000)
    ap.add_argument("--pose", default="default", choices=["default", "pinned"])
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the last step")
    ap.add_argument("--no-c4", action="store_true",
                    help="skip the secondary records (C4: 3840x2160 x16; C3: Cornell box 1080p x16)")
    ap.add_argument("--frame-by-frame", action="store_true",
                    help="one rvpt_b200_render_frame call (= one launch) per frame instead of one "
                         "rvpt_b200_render_frames batch per step")
    ap.add_argument("--graph", default="off", choices=["on", "off"],
                    help="replay each step from a CUDA graph (captured through the C ABI)")
    return ap.parse_args()


def workload_name(args) -> str:
    if args.scene in ("builtin", "cornell"):
        return f"{args.scene}_{args.pose if args.scene == 'builtin' else 'default'}"
    return args.scene


def describe(args, n_tris, n_nodes, W, H, frames) -> str:
    cfg = {"builtin": "C2", "cornell": "C3"}.get(args.scene, "f-1")
    if (W, H) == (3840, 2160):
        cfg = "C4"
    return (f"{cfg}: {args.scene} scene ({n_tris} triangles, {n_nodes} BVH nodes), {args.pose} pose, "
            f"{W}x{H}, max_bounces {args.bounces}, aa {args.aa}, progressive frames 0..{frames - 1}")


def committed_workload(args):
    """nodes, triangles (BVH order), materials, camera for the 16:9 configurations — from
    oracle/workloads/ (no product code involved). None for generated scenes."""
    import oracle
    name = workload_name(args)
    if not (ROOT / "oracle" / "workloads" / f"{name}.npz").exists():
        return None
    if abs(args.width / args.height - 16 / 9) > 1e-9:
        return None
    w = oracle.load_workload(name)
    return w["nodes"], w["triangles"], w["materials"], w["camera_16x9"]


def product_workload(args):
    """The same arrays built at run time with the product's host helpers (GPU arm)."""
    import rvpt_b200 as rv
    if args.scene == "builtin":
        scene = rv.builtin_scene()
        pose, fov = ((0.0, 0.0, 0.0) if args.pose == "default" else (0.0, 0.8, -2.5)), 90.0
    elif args.scene == "mesh":
        scene = rv.displaced_sphere_scene(args.mesh_tris)
        pose, fov = (0.0, 1.2, -3.0), 60.0
    elif args.scene == "tridel":
        from rvpt_b200.scene import tridel_scene
        scene, pose, fov = tridel_scene()
    else:
        scene = rv.cornell_scene()
        pose, fov = (0.0, 1.2, -3.4), 60.0
    nodes, perm = rv.build_bvh(scene.triangles)
    tris = np.ascontiguousarray(scene.triangles[perm])
    cam = rv.camera_data(translation=pose, aspect=args.width / args.height, fov=fov)
    return nodes, tris, scene.materials, cam


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region. Polls NVML (the library behind
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`) every 2 ms from a
    thread between start() and stop()."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[float, int]] = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001
            self.err = f"NVML unavailable: {exc}"
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except AttributeError:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(mask)))
            except Exception as exc:  # noqa: BLE001
                self.err = str(exc)
                break
            time.sleep(0.002)

    def stop(self) -> dict:
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "n/a"]}
        self.stop_flag.set()
        self.thread.join(timeout=1)
        sm = [m for m, _ in self.samples]
        reasons = set()
        for _, mask in self.samples:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_renderer(W, H, nodes, tris, mats):
    """The CPU implementation of the path that serves as checker and baseline. kind "reference":
    the reference's own shipped shader binary translated to C++ and compiled for the host
    (oracle/_ref/libref_shader.so, built next to the reference tree by `make -C oracle ref_shader`;
    the prebuilt library travels to the GPU box). kind "port": the restatement of the GLSL,
    oracle/rvpt_oracle.cpp, when that library does not exist. Returns (renderer, kind, cores)."""
    import oracle
    try:
        from oracle import ref_shader
        if ref_shader.LIB_PATH.exists() or ref_shader.REFERENCE_SPV.exists():
            r = ref_shader.RefShaderRenderer(W, H, tris, mats, nodes)
            return r, "reference", r.threads()
    except Exception as exc:  # noqa: BLE001
        print(f"bench: reference shader library unavailable ({exc}); using the oracle port", file=sys.stderr)
    ora = oracle.OracleRenderer(W, H, tris, mats, nodes)
    return ora, "port", oracle.load().rvpt_oracle_hardware_threads()


def cpu_image(r, kind):
    return r.result_rgba8() if kind == "reference" else r.result


def oracle_frames(W, H, nodes, tris, mats, cam, frames, bounces, aa):
    """Renders frames 0..frames-1 on the host CPU. Returns (rgba8 image, seconds, cores, kind)."""
    import oracle
    r, kind, cores = cpu_renderer(W, H, nodes, tris, mats)
    t = time.perf_counter()
    for f in range(frames):
        r.render_frame(oracle.settings(max_bounces=bounces, aa=aa, frame=f), cam)
    return cpu_image(r, kind), time.perf_counter() - t, cores, kind


def run_reference(args):
    """Reference arm: the reference's own implementation of the path on the box's host cores — its
    shipped shader binary compiled for the CPU (kind "reference") or, without that library, the
    restatement of the GLSL (kind "port"). The Vulkan path itself cannot run here. Loads oracle/ only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    from oracle import vulkan_probe
    wl = committed_workload(args)
    if wl is None:  # generated scenes have no committed copy: build them with the host helpers
        wl = product_workload(args)
    nodes, tris, mats, cam = wl
    W, H = args.width, args.height
    ora, kind, cores = cpu_renderer(W, H, nodes, tris, mats)
    frame = 0

    def step():  # one full frame of the workload (a bounded sample of the 64-frame step)
        nonlocal frame
        ora.render_frame(oracle.settings(max_bounces=args.bounces, aa=args.aa, frame=frame % args.frames), cam)
        frame += 1

    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t
    msps = args.steps * W * H * args.aa / dt / 1e6
    sample = (f"each step = 1 full frame ({W}x{H}x{args.aa} samples) of the workload's {args.frames}-frame "
              f"progressive batch, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": msps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": describe(args, len(tris), len(nodes), W, H, args.frames),
                   "frames_per_step": args.frames},
        "cpu_baseline": {"value": msps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "what": {"reference": "the reference's shipped compute_pass.comp.spv translated to C++ and "
                                               "compiled for the host (oracle/_ref/libref_shader.so)",
                                  "port": "CPU restatement of the GLSL (oracle/rvpt_oracle.cpp)"}[kind],
                         # could the reference's Vulkan path itself run here (a software ICD)? SURVEY 8 f-4
                         "vulkan": vulkan_probe.probe()},
        "e2e": {"value": msps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class Runner:
    """One engine + the per-step plumbing of one configuration (resolution, frames per step)."""

    def __init__(self, args, torch, dist, rv, wl, W, H, F, rank, world, local_rank, stream):
        self.args, self.torch, self.dist, self.rv = args, torch, dist, rv
        self.W, self.H, self.F, self.rank, self.world = W, H, F, rank, world
        self.nodes, self.tris, self.mats, self.cam = wl
        self.eng = rv.Engine(W, H, device=local_rank, rank=rank, nranks=world)
        self.eng.set_stream(stream.cuda_stream)
        self.eng.upload_scene(self.tris, self.mats, self.nodes)
        self.settings = [rv.default_settings(max_bounces=args.bounces, aa=args.aa, frame=f) for f in range(F)]
        self.settings_ptr = [s.ctypes.data for s in self.settings]
        self.cam_ptr = self.cam.ctypes.data
        self.fg = self.po = None
        self.gather = args.gather if world > 1 else "none"
        dev = torch.device("cuda", local_rank)
        if self.gather == "nccl":
            from rvpt_b200.distributed import FrameGather
            self.fg = FrameGather(self.eng, dist, torch, dev)
        elif self.gather == "p2p":
            from rvpt_b200.distributed import PeerOutput
            self.po = PeerOutput(self.eng, dist, torch, dev)
        self.out_pinned = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)
        self.out_np = self.out_pinned.numpy()
        self.out_pinned2 = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)  # e2e double buffer
        self.e2e_bufs = [self.out_np, self.out_pinned2.numpy()]
        self.e2e_k = 0

    def frames_of_step(self):
        if self.fg:
            self.fg.begin_frame()  # device-side wait for the buffer's previous gather
        if self.args.frame_by_frame:
            for f in range(self.F):
                self.eng.render_frame_raw(self.settings_ptr[f], self.cam_ptr)
        else:
            self.eng.render_frame_raw(self.settings_ptr[0], self.cam_ptr, self.F)
        if self.fg:
            self.fg.end_frame()    # ONE all-gather for the step's image + untile on rank 0
            self.fg.flush()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def read_image(self):
        """The step's assembled rgba8 image on rank 0 (device -> pinned host)."""
        if self.world == 1:
            self.eng.read_output_rgba8(self.out_np)
        elif self.po:
            self.po.finish()       # every rank's pixels have landed in rank 0's image
            if self.rank == 0:
                self.eng.read_output_rgba8(self.out_np)
        elif self.fg:
            if self.rank == 0:
                self.out_pinned.view(self.torch.int32).view(-1).copy_(self.fg.raster, non_blocking=False)
            else:
                self.torch.cuda.synchronize()
        else:
            self.torch.cuda.synchronize()
        return self.out_np

    def e2e_step(self):
        """One end-to-end step: scene from host arrays, the frames, the image back into pinned host
        memory. The read-back is double-buffered (rvpt_b200_read_output_rgba8_async): the copy of
        step k overlaps the frames of step k + 1; e2e_finish() waits for the last one."""
        self.eng.upload_scene(self.tris, self.mats, self.nodes)  # host arrays -> device, every step
        self.frames_of_step()
        buf = self.e2e_bufs[self.e2e_k & 1]
        self.e2e_k += 1
        if self.world == 1:
            self.eng.wait_output()                     # the copy that used `buf` two steps ago
            self.eng.read_output_rgba8_async(buf)
        elif self.po:
            self.po.read_image_async(buf)
        else:
            self.read_image()

    def e2e_finish(self):
        if self.rank == 0 and (self.world == 1 or self.po):
            self.eng.wait_output()

    def time_steps(self, steps, warmup, stream, flush, graph):
        """Device-timed steps: returns (total ms as max over ranks, launches per step, stats)."""
        torch, dist = self.torch, self.dist
        for _ in range(max(warmup, 3)):
            self.frames_of_step()
        self.barrier()
        step_graph, note = None, "direct launches"
        if graph == "on" and self.gather != "nccl":
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                self.frames_of_step()
            step_graph, note = g, "one CUDA graph per step (captured through the C ABI)"
            step_graph.replay()
            self.barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        self.barrier()
        for k in range(steps):
            flush.zero_()          # evict the accumulation image and the queues from L2 between steps
            if self.world > 1:
                dist.barrier()
            starts[k].record(stream)
            if step_graph is not None:
                step_graph.replay()
            else:
                self.frames_of_step()
            ends[k].record(stream)
        self.barrier()
        total = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], dtype=torch.float64,
                             device="cuda")
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), note

    def parity(self):
        """The image of the last step (already rendered) against the oracle's frames 0..F-1."""
        img = self.read_image()
        ok = None
        if self.rank == 0:
            want, secs, cores, kind = oracle_frames(self.W, self.H, self.nodes, self.tris, self.mats, self.cam,
                                                    self.F, self.args.bounces, self.args.aa)
            ok = bool(np.array_equal(img, want))
            self.oracle_run = (secs, cores, kind)
            self.image_sha256 = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
            self.diff_pixels = int((img != want).any(axis=-1).sum())
        if self.world > 1:
            self.dist.barrier()
        return ok


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import rvpt_b200 as rv
    wl = product_workload(args)
    committed = committed_workload(args)
    if committed is not None:  # both arms and the parity check see the same bytes
        for a, b in zip(wl, committed):
            assert np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes(), \
                "oracle/workloads is stale: run tools/make_bench_workloads.py"
    nodes, tris, mats, cam = wl
    W, H, F = args.width, args.height, args.frames
    # a real (non-default) stream shared by torch and the engine, so torch's CUDA events bracket
    # exactly the kernels the C ABI launches
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # 2x L2

    run = Runner(args, torch, dist, rv, wl, W, H, F, rank, world, local_rank, stream)
    eng = run.eng

    # ---- device-resident throughput --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    # warm-up + graph capture happen inside time_steps before its timed loop; the sampler is
    # started right before and only keeps what it saw under load (median)
    if rank == 0:
        sampler.start()
    total_ms, launch_note = run.time_steps(args.steps, args.warmup, stream, flush, args.graph)
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()                       # counters of the last launch (st["frames"] frames)
    launches_per_step = st["kernel_launches"] + (1 if (run.fg and rank == 0) else 0)
    samples_per_step = W * H * args.aa * F
    value = args.steps * samples_per_step / (total_ms * 1e-3) / 1e6

    # ---- parity of what was just timed -------------------------------------------------------
    parity_ok = None if args.no_parity else run.parity()

    # ---- roofline of the dominant kernel (separate pass, CUDA events around every launch) -----
    eng.set_profiling(True)
    eng.kernel_times()
    for _ in range(3):
        flush.zero_()
        run.frames_of_step()
    kt = eng.kernel_times()
    eng.set_profiling(False)
    active = st["active"] + [0] * 64
    n_f = max(st["frames"], 1)             # frames the last launch covered
    S = st["samples"]                      # samples of that launch (this rank)
    R = sum(active)
    P = S // (n_f * args.aa) if n_f else 0  # pixels of this rank
    # DESIGN.md "algorithmic bytes": every segment after the first writes and re-reads the 64 B
    # path state (128 B). Batched launch (n_f > 1): a finished sample is parked (16 B written,
    # 16 B read back by the resolve phase) and each pixel's running mean + rgba8 pixel move once
    # per launch (36 B). Frame-by-frame launch: 36 B per sample.
    if n_f > 1:
        launch_bytes = 32 * S + 36 * P + 128 * (R - S)
    else:
        launch_bytes = 36 * S + 128 * (R - S)
    survey_bytes = 128 * R + 36 * S        # SURVEY 8(d) literal formula (generate/trace unfused)
    kernel_ms = kt["primary_ms"] / max(kt["primary_launches"], 1)
    peak, peak_src = measured_hbm_peak()
    achieved = launch_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    traffic = None
    tpath = ROOT / "profiles" / "k_frame_traffic.json"
    default_wl = (args.scene, args.pose, W, H, args.bounces, args.aa, world, F, args.frame_by_frame) == \
        ("builtin", "default", 1920, 1080, 8, 1, 1, 64, False)
    if tpath.exists() and default_wl:
        tj = json.loads(tpath.read_text())
        if tj.get("frames_per_launch") == n_f:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {
        "bound": "hbm",
        "kernel": f"k_frame<batched> (one persistent cooperative launch per {n_f} frames: merged primary + "
                  "bounce waves, parked samples, in-order resolve)" if n_f > 1 else
                  "k_frame (one persistent cooperative launch per frame)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": traffic, "bytes_per_launch": launch_bytes, "ms_per_launch": kernel_ms,
        "frames_per_launch": n_f, "launches_timed": kt["primary_launches"],
        "survey_8d_formula_bytes": survey_bytes,
        "survey_8d_formula_gbs": survey_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0,
        "rays_per_sample": R / max(S, 1), "active_per_bounce_last_launch": st["active"],
        "note": "the kernel is FP32-issue bound, not HBM bound (DESIGN.md section 4); L2 is flushed between steps",
    }

    # ---- end to end through the host-buffer API ------------------------------------------------
    h2d = nodes.nbytes + tris.nbytes + mats.nbytes + 40 + 80
    d2h = W * H * 4
    for _ in range(3):
        run.e2e_step()
    run.e2e_finish()
    run.barrier()
    e2e_steps = max(5, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        run.e2e_step()
    run.e2e_finish()
    run.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * samples_per_step / float(e2e_s.item()) / 1e6

    # ---- CPU baseline (rank 0, N = 1 only): the oracle run of the parity check, timed ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if getattr(run, "oracle_run", None) is None:
            _, secs, cores, kind = oracle_frames(W, H, nodes, tris, mats, cam, min(F, 16), args.bounces, args.aa)
            n_frames = min(F, 16)
        else:
            (secs, cores, kind), n_frames = run.oracle_run, F
        cpu = {"value": n_frames * W * H * args.aa / secs / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_frames} full frames of the workload (the CPU run of the parity check), "
                         f"{secs:.1f} s on {cores} threads",
               "what": {"reference": "the reference's shipped compute_pass.comp.spv translated to C++ and compiled "
                                     "for the host (oracle/_ref/libref_shader.so)",
                        "port": "CPU restatement of the GLSL (oracle/rvpt_oracle.cpp)"}[kind]}
        if kind == "reference":
            # for comparison: the hand-written restatement of the GLSL on the same cores (4 frames)
            import oracle
            ora = oracle.OracleRenderer(W, H, tris, mats, nodes)
            t0p = time.perf_counter()
            for f in range(4):
                ora.render_frame(oracle.settings(max_bounces=args.bounces, aa=args.aa, frame=f), cam)
            cpu["port_value"] = 4 * W * H * args.aa / (time.perf_counter() - t0p) / 1e6

    # ---- secondary records: C4 (3840x2160 x 16 frames) and C3 (Cornell box, 1080p x 16 frames) ----
    c4 = c3 = None
    if not args.no_c4 and args.scene == "builtin" and (W, H) == (1920, 1080):
        run.barrier()
        run.eng.close()

        def secondary(sargs, swl, sw, sh, sframes):
            r = Runner(sargs, torch, dist, rv, swl, sw, sh, sframes, rank, world, local_rank, stream)
            ssteps = max(5, args.steps // 2)
            ms, _ = r.time_steps(ssteps, args.warmup, stream, flush, "off")
            ok = None if args.no_parity else r.parity()
            st2 = r.eng.stats()
            rec = None
            if rank == 0:
                rec = {"workload": describe(sargs, len(swl[1]), len(swl[0]), sw, sh, sframes),
                       "value": ssteps * sw * sh * sframes * args.aa / (ms * 1e-3) / 1e6, "unit": UNIT,
                       "ms_per_step": ms / ssteps, "steps": ssteps, "frames_per_step": sframes,
                       "launches_per_step": st2["kernel_launches"], "parity_ok": ok}
            r.barrier()
            r.eng.close()
            return rec

        c4_args = argparse.Namespace(**vars(args))
        c4_args.width, c4_args.height, c4_args.frames = 3840, 2160, 16
        c4 = secondary(c4_args, wl, 3840, 2160, 16)
        c3_args = argparse.Namespace(**vars(args))
        c3_args.scene, c3_args.frames = "cornell", 16
        c3_wl = product_workload(c3_args)
        c3 = secondary(c3_args, c3_wl, 1920, 1080, 16)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": describe(args, len(tris), len(nodes), W, H, F), "frames_per_step": F},
            "run": {"samples_per_step": samples_per_step,
                    "partition": f"16x16 tiles, tile_id % {world} == rank" if world > 1 else "none",
                    "gather": {"p2p": "fused: the kernels store rgba8 pixels into rank 0's raster image over "
                                      "NVLink peer memory (CUDA IPC); one barrier per step",
                               "nccl": "one rgba8 all_gather_into_tensor per step + untile on rank 0",
                               "none": "none"}[run.gather],
                    "launch": ("one rvpt_b200_render_frame call per frame; " if args.frame_by_frame else
                               f"one rvpt_b200_render_frames({F}) call per step = {st['kernel_launches']} batched "
                               f"launch(es) of <= {n_f} frames; ") + launch_note,
                    "l2": "256 MiB memset between steps (outside the timed events); inputs of a step "
                          "(accumulation image, queues, parked samples: > 126 MB L2 per launch) stream from HBM"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "what": "upload_scene(host arrays) + render_frames + read_output_rgba8_async into pinned "
                            "memory (double-buffered: the copy of step k overlaps the frames of step k+1; the "
                            "last copy is waited for inside the timed region), wall clock incl. synchronisation"},
            "gpu_launches": args.steps * launches_per_step,
            "parity_ok": parity_ok,
            "parity": None if args.no_parity else {
                "what": f"rgba8 image after the last timed step vs frames 0..{F - 1} rendered on the CPU by "
                        + ("the reference's own compiled shader" if getattr(run, "oracle_run", (0, 0, "port"))[2] == "reference"
                           else "the oracle"),
                "differing_pixels": getattr(run, "diff_pixels", None),
                "image_sha256": getattr(run, "image_sha256", None)},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "c4": c4, "c3": c3,
            "mrays_per_s": value * (R / max(S, 1)),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
