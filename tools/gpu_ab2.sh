#!/bin/bash
# A/B of library variants (tools/build_variant.sh): bench lines for builtin + cornell (+ mesh with "mesh")
mkdir -p gpurun_out
TAG=${1:-ab}; shift
for v in "$@"; do
  for wl in "builtin:--steps 10" "cornell:--steps 3 --scene cornell" ; do
    name=${wl%%:*}; args=${wl#*:}
    RVPT_B200_LIB=$PWD/rvpt_b200/variants/lib$v.so timeout 300 python bench.py --no-cpu-baseline --no-c4 --no-parity $args > gpurun_out/bench_${TAG}_${v}_${name}.json 2> gpurun_out/bench_${TAG}_${v}_${name}.err
    python - "$TAG" "$v" "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s_%s_%s.json" % tuple(sys.argv[1:4])))
    r = d["roofline"]
    print(sys.argv[2], sys.argv[3], "value", round(d["value"]), "ms/launch", round(r["ms_per_launch"], 3), "frames/launch", r["frames_per_launch"])
except Exception as e:
    print(sys.argv[2], sys.argv[3], "failed", e)
PY
  done
done
