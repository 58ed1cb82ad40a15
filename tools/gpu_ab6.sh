#!/bin/bash
mkdir -p gpurun_out
run() { # name flags args
  local name=$1 fl=$2; shift 2
  RVPT_B200_EXTRA_FLAGS=$fl timeout 600 python bench.py --no-cpu-baseline --no-c4 "$@" > gpurun_out/bench_ab6_${name}.json 2> gpurun_out/bench_ab6_${name}.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab6_%s.json" % sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], d["parity"] and d["parity"]["differing_pixels"], "ms/launch", round(r["ms_per_launch"], 3))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run mesh500k 0 --scene mesh --mesh-tris 500000 --frames 16 --steps 3
run mesh500k_ordered 0x800 --scene mesh --mesh-tris 500000 --frames 16 --steps 3
run tridel 0 --scene tridel --frames 16 --steps 3
run tridel_ordered 0x800 --scene tridel --frames 16 --steps 3
run mesh20k_ordered 0x800 --scene mesh --mesh-tris 20000 --frames 16 --steps 3
tail -3 gpurun_out/bench_ab6_tridel_ordered.err
