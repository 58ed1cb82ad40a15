#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # name flags args
  local name=$1 fl=$2; shift 2
  RVPT_B200_EXTRA_FLAGS=$fl timeout 600 python bench.py --no-cpu-baseline --no-c4 "$@" > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err
  python - "$TAG" "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], sys.argv[2])))
    r = d["roofline"]
    print(sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "ms/launch", round(r["ms_per_launch"], 3), "frames/launch", r["frames_per_launch"], r["active_per_bounce_last_launch"][:4])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run mesh500k 0 --scene mesh --mesh-tris 500000 --frames 16 --steps 3
run mesh20k 0 --scene mesh --mesh-tris 20000 --frames 16 --steps 3
run tridel 0 --scene tridel --frames 16 --steps 3
