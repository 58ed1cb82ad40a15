#!/bin/bash
# leaf server A/B: the closed-scene tests, then bench lines with and without it. usage: gpu_srv.sh <tag>
mkdir -p gpurun_out
TAG=${1:-srv}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "leaf_server or cornell or stated_configs or batched" 2>&1 | tail -5
for v in on off; do
  F=0; [ $v = off ] && F=0x800
  RVPT_B200_AB_FLAGS=$F timeout 600 python bench.py --scene cornell --steps 5 --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_cornell_$v.json 2>gpurun_out/bench_${TAG}_cornell_$v.err
  RVPT_B200_AB_FLAGS=$F timeout 600 python bench.py --no-cpu-baseline --no-c4 > gpurun_out/bench_${TAG}_builtin_$v.json 2>gpurun_out/bench_${TAG}_builtin_$v.err
done
python - "$TAG" <<'PY'
import json, sys
for n in ("cornell_on", "cornell_off", "builtin_on", "builtin_off"):
    try:
        d = json.loads([l for l in open("gpurun_out/bench_%s_%s.json" % (sys.argv[1], n)) if l.startswith("{")][-1])
        print(n, "value", round(d["value"]), "parity", d["parity_ok"], "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/bench_${TAG}_cornell_on.err
