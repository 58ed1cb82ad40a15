#!/bin/bash
# A/B of library variants (tools/build_variant.sh) on ONE box (boxes differ by 3-4 %).
# usage: gpu_ab_variants.sh <tag> "<workload> ..." <variant...>
#   workload: builtin | pinned | cornell | tridel | mesh500k ; variant "default" = the in-tree library
mkdir -p gpurun_out
TAG=$1; WL=$2; shift 2
for v in "$@"; do
  for w in $WL; do
    case $w in
      builtin) A="";; pinned) A="--pose pinned";; cornell) A="--scene cornell --steps 5";;
      tridel) A="--scene tridel --frames 16 --steps 3";; mesh500k) A="--scene mesh --mesh-tris 500000 --frames 16 --steps 3";;
    esac
    L=$PWD/rvpt_b200/variants/lib$v.so; [ $v = default ] && L=
    RVPT_B200_LIB=$L timeout 600 python bench.py $A --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_${w}_$v.json 2>gpurun_out/bench_${TAG}_${w}_$v.err
  done
done
python - "$TAG" "$WL" "$@" <<'PY'
import json, sys
for v in sys.argv[3:]:
    for w in sys.argv[2].split():
        try:
            d = json.loads([l for l in open("gpurun_out/bench_%s_%s_%s.json" % (sys.argv[1], w, v)) if l.startswith("{")][-1])
            print(v, w, "value", round(d["value"]), "parity", d["parity_ok"], "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
        except Exception as e:
            print(v, w, "failed", e)
PY
