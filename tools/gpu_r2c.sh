#!/bin/bash
# parity tests + the three workloads, default vs NO_QUEUE_SORT (0x100)
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run() { # name, extra flags env, args...
  local name=$1 fl=$2; shift 2
  RVPT_B200_EXTRA_FLAGS=$fl timeout 400 python bench.py --no-cpu-baseline --no-c4 "$@" > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err
  python - "$TAG" "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], sys.argv[2])))
    r = d["roofline"]
    print(sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "ms/launch",
          round(r["ms_per_launch"], 3), "frames/launch", r["frames_per_launch"], r["active_per_bounce_last_launch"][:4])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run builtin 0 --steps 20
run pinned 0 --steps 20 --pose pinned
run cornell 0 --steps 5 --scene cornell
run cornell_nosort 0x100 --steps 5 --scene cornell --no-parity
run builtin_nosort 0x100 --steps 20 --no-parity
timeout 200 python tools/timeline.py --batch 32 --scene cornell > gpurun_out/timeline_${TAG}_cornell_b32.md 2>&1
cat gpurun_out/timeline_${TAG}_cornell_b32.md
tail -3 gpurun_out/bench_${TAG}_cornell.err
