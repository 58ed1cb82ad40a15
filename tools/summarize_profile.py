"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the
small tracked summaries under profiles/.

    python tools/summarize_profile.py <tag>     # e.g. r01

Inputs : gpurun_out/launches_<tag>.csv, gpurun_out/prof_frame_<tag>.ncu-rep,
         gpurun_out/bench_<tag>_*.json, gpurun_out/timeline_<tag>_*.md, gpurun_out/<tag>_sanitizer_*.log
Outputs: profiles/<tag>_launches.md, profiles/<tag>_k_frame_metrics.md,
         profiles/<tag>_k_frame_hot_blocks.txt, profiles/k_frame_traffic.json, copies of the bench lines,
         timelines and sanitizer summaries
"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
PROFILE_CMD = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-parity"
PROF.mkdir(exist_ok=True)

# ---- launch list -----------------------------------------------------------------
lines = [l for l in open(OUT / f"launches_{tag}.csv") if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
    agg.setdefault(name, []).append(v)
total = sum(sum(v) for v in agg.values())
with open(PROF / f"{tag}_launches.md", "w") as f:
    f.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
            f"`{PROFILE_CMD}` (64-frame steps: one batched `k_frame` launch each; the other kernels are "
            "torch's L2-flush memset and the roofline / e2e passes of bench.py).\n"
            "Per-launch times are cold-cache and serialised: compare shares, not absolutes.\n\n"
            "| kernel | launches | mean us | min us | max us | share of device time |\n|---|---|---|---|---|---|\n")
    for k, v in agg.items():
        f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {100*sum(v)/total:.1f}% |\n")

# ---- k_frame raw metrics -----------------------------------------------------------
rep = OUT / f"prof_frame_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
]
with open(PROF / f"{tag}_k_frame_metrics.md", "w") as f:
    f.write(f"# k_frame (batched, 64 frames per launch), `ncu --set full --clock-control none` ({tag})\n\n"
            "Workload: C2 (built-in scene, default pose, 1920x1080, 8 bounces, aa 1, frames 0..63 in one launch). "
            "One column per captured launch.\n\n"
            "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(vals))) + " |\n|---|---|" + "---|" * len(vals) + "\n")
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            f.write(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in vals) + " |\n")
    f.write("\n## warp stall reasons (warps per issue-active cycle)\n\n| reason | value |\n|---|---|\n")
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                if float(vals[0][i]) > 0.05:
                    f.write(f"| `{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}` | {float(vals[0][i]):.2f} |\n")
            except ValueError:
                pass

def _bytes(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return [float(r[i]) * scale for r in vals]


rd, wr = _bytes("dram__bytes_read.sum"), _bytes("dram__bytes_write.sum")
(PROF / "k_frame_traffic.json").write_text(json.dumps({
    "tag": tag, "kernel": "k_frame<batched>", "launches_captured": len(vals), "frames_per_launch": 64,
    "dram_bytes_per_launch": sum(rd + wr) / len(vals),
    "dram_bytes_read": rd, "dram_bytes_write": wr,
    "source": f"ncu --set full --clock-control none, gpurun_out/prof_frame_{tag}.ncu-rep "
              f"({PROFILE_CMD})"}, indent=1) + "\n")

src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
(OUT / f"src_{tag}.csv").write_text(src)
blocks = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_blocks.py"), str(OUT / f"src_{tag}.csv"), "1.0"],
                        capture_output=True, text=True).stdout
(PROF / f"{tag}_k_frame_hot_blocks.txt").write_text(
    "SASS basic-block groups of k_frame with >= 1% of executed warp-instructions\n"
    "(start-end SASS index, instructions in block, executions, share of instructions, share of stall samples,\n"
    " avg active threads, top opcodes). The 27-instruction FMNMX/FMUL/LDS blocks are the BVH node slab test.\n\n" + blocks)
for src in sorted(OUT.glob(f"bench_{tag}_*.json")):
    lines = [l for l in src.read_text().splitlines() if l.startswith("{")]
    if lines:
        (PROF / src.name.replace(f"bench_{tag}_", f"{tag}_bench_")).write_text(lines[-1] + "\n")
for src in sorted(OUT.glob(f"timeline_{tag}_*.md")):
    shutil.copy(src, PROF / src.name.replace(f"timeline_{tag}_", f"{tag}_timeline_"))
for src in sorted(OUT.glob(f"{tag}_sanitizer_*.log")):
    text = src.read_text().splitlines()
    keep = [l for l in text if "SUMMARY" in l or "ERROR" in l or "smoke ok" in l or "batch ok" in l][-12:]
    (PROF / src.name).write_text("\n".join(keep) + "\n")
print("wrote", sorted(p.name for p in PROF.glob(f"{tag}_*")))
