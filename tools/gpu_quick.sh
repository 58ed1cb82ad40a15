#!/bin/bash
# quick loop on one B200: all GPU tests + the three bench workloads with the default flags (no ncu)
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|words differ|Error" | tail -12
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
timeout 300 python bench.py --steps 30 --warmup 3 --pose pinned --no-cpu-baseline > gpurun_out/bench_${TAG}_pinned.json 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --scene cornell --no-cpu-baseline > gpurun_out/bench_${TAG}_cornell.json 2>/dev/null
python - "$TAG" <<'PY'
import json, sys
for n in ("n1", "pinned", "cornell"):
    try:
        d = json.load(open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], n)))
        r = d["roofline"]
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/frame",
              round(r["frame_ms_in_timed_region"] * 1000, 1), "kernel us", round(r["ms_per_launch"] * 1000, 1), "frac",
              round(r["frac"], 3), r["active_per_bounce"][:4], d["clocks"]["sm_mhz"], d.get("cpu_baseline"))
    except Exception as e:
        print(n, "failed", e)
PY
