#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  for wl in "mesh500k:--scene mesh --mesh-tris 500000 --frames 16 --steps 3" "tridel:--scene tridel --frames 16 --steps 3" "mesh20k:--scene mesh --mesh-tris 20000 --frames 16 --steps 3"; do
    name=${wl%%:*}; args=${wl#*:}
    RVPT_B200_LIB=$PWD/rvpt_b200/variants/lib$v.so timeout 400 python bench.py --no-cpu-baseline --no-c4 $args > gpurun_out/bench_ab5_${v}_${name}.json 2> gpurun_out/bench_ab5_${v}_${name}.err
    python - "$v" "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab5_%s_%s.json" % tuple(sys.argv[1:3])))
    print(sys.argv[1], sys.argv[2], "value", round(d["value"]), "parity", d["parity_ok"], "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e)
PY
  done
done
