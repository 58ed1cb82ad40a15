#!/bin/bash
# round 2, first GPU pass: parity tests, then batched vs frame-by-frame launches on the three workloads
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
run() { # name, extra flags env, args...
  local name=$1 fl=$2; shift 2
  RVPT_B200_EXTRA_FLAGS=$fl timeout 400 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err
  python - "$TAG" "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], sys.argv[2])))
    r = d["roofline"]
    print(sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/frame",
          round(r["frame_ms_in_timed_region"] * 1000, 2), "kernel us", round(r["ms_per_launch"] * 1000, 1),
          r["active_per_bounce"][:4], d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run batch16 0 --steps 30 --warmup 3
run nobatch16 0x200 --steps 30 --warmup 3
run batch64 0 --steps 10 --warmup 3 --frames 64
run pinned_batch16 0 --steps 30 --warmup 3 --pose pinned
run pinned_nobatch16 0x200 --steps 30 --warmup 3 --pose pinned
run cornell_batch16 0 --steps 10 --warmup 3 --scene cornell
run cornell_nobatch16 0x200 --steps 10 --warmup 3 --scene cornell
tail -3 gpurun_out/bench_${TAG}_batch16.err
