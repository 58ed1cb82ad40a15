#!/bin/bash
# A/B of two engine flag sets on one B200: all GPU tests, then the three bench workloads per variant, then timelines.
#   gpurun -- 'bash tools/gpu_ab.sh 0x0 0x80'      (default vs RVPT_B200_FLAG_REFERENCE_ORDER)
# Flags are OR-ed into every Engine through RVPT_B200_EXTRA_FLAGS (rvpt_b200/engine.py); see include/rvpt_abi.h.
A=${1:-0x0}; B=${2:-0x80}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|words differ|Error" | tail -12
for v in $A $B; do
  export RVPT_B200_EXTRA_FLAGS=$v
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${v}_n1.json 2> gpurun_out/bench_${v}_n1.err
  timeout 300 python bench.py --steps 30 --warmup 3 --pose pinned --no-cpu-baseline > gpurun_out/bench_${v}_pinned.json 2>/dev/null
  timeout 300 python bench.py --steps 10 --warmup 3 --scene cornell --no-cpu-baseline > gpurun_out/bench_${v}_cornell.json 2>/dev/null
done
unset RVPT_B200_EXTRA_FLAGS
python - "$A" "$B" <<'PY'
import json, sys
for v in sys.argv[1:3]:
    for n in ("n1", "pinned", "cornell"):
        try:
            d = json.load(open("gpurun_out/bench_%s_%s.json" % (v, n)))
            r = d["roofline"]
            print("flags", v, n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/frame",
                  round(r["frame_ms_in_timed_region"] * 1000, 1), "frac", round(r["frac"], 3),
                  r["active_per_bounce"][:4], d["clocks"]["sm_mhz"])
        except Exception as e:
            print(v, n, "failed", e)
PY
timeout 120 python tools/timeline.py > gpurun_out/timeline_builtin.md 2>&1; tail -8 gpurun_out/timeline_builtin.md
timeout 120 python tools/timeline.py --scene cornell > gpurun_out/timeline_cornell.md 2>&1; tail -17 gpurun_out/timeline_cornell.md
