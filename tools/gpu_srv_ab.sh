#!/bin/bash
# leaf server variants (tools/build_variant.sh) on one box. usage: gpu_srv_ab.sh <tag> <variant...>
mkdir -p gpurun_out
TAG=$1; shift
for v in "$@"; do
  for sc in cornell builtin; do
    RVPT_B200_LIB=$PWD/rvpt_b200/variants/lib$v.so timeout 600 python bench.py --scene $sc --steps 5 --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_${sc}_$v.json 2>gpurun_out/bench_${TAG}_${sc}_$v.err
  done
done
python - "$TAG" "$@" <<'PY'
import json, sys
for v in sys.argv[2:]:
    for sc in ("cornell", "builtin"):
        try:
            d = json.loads([l for l in open("gpurun_out/bench_%s_%s_%s.json" % (sys.argv[1], sc, v)) if l.startswith("{")][-1])
            print(v, sc, "value", round(d["value"]), "parity", d["parity_ok"], "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
        except Exception as e:
            print(v, sc, "failed", e)
PY
