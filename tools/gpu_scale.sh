#!/bin/bash
# bench.py at N GPUs of one box, launched the way the driver launches it. usage: gpu_scale.sh <tag> <N...>
mkdir -p gpurun_out
TAG=${1:-scale}; shift
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
  fi
  python - "$TAG" "$N" <<'PY'
import json, sys
try:
    lines = [l for l in open("gpurun_out/bench_%s_n%s.json" % (sys.argv[1], sys.argv[2])) if l.startswith("{")]
    d = json.loads(lines[-1])
    print("N", sys.argv[2], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "ms/step", round(d["ms_per_step"], 3),
          "c4", d["c4"] and (round(d["c4"]["value"]), d["c4"]["parity_ok"]), "c3", d.get("c3") and (round(d["c3"]["value"]), d["c3"]["parity_ok"]), "launches", d["gpu_launches"], d["run"]["launch"][:60])
except Exception as e:
    print("N", sys.argv[2], "failed", e)
PY
  tail -2 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
done
