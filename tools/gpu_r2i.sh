#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # name flags args
  local name=$1 fl=$2; shift 2
  RVPT_B200_EXTRA_FLAGS=$fl timeout 600 python bench.py --no-cpu-baseline --no-c4 "$@" > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err
  python - "$TAG" "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], sys.argv[2])))
    r = d["roofline"]
    print(sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "ms/launch", round(r["ms_per_launch"], 3), "frames/launch", r["frames_per_launch"])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run cornell 0 --steps 5 --scene cornell
run cornell_nodefer 0x400 --steps 5 --scene cornell --no-parity
run builtin 0 --steps 10 --no-parity
run builtin_nodefer 0x400 --steps 10 --no-parity
for fl in 0 0x400; do
  echo "== timeline cornell flags $fl"
  timeout 200 python tools/timeline.py --batch 16 --scene cornell --flags $fl 2>&1 | grep -E "^\| (2|4|6|8|10|12|14|15) "
done
