#!/bin/bash
# which source change makes C2 slow (25 instead of 30 Gsamples/s)? five builds of kernels.cu on one box, + the phase timeline of the good and the bad one
mkdir -p gpurun_out
V=$PWD/rvpt_b200/variants
for v in good opq syn loc bad; do
  RVPT_B200_LIB=$V/lib$v.so timeout 300 python bench.py --no-cpu-baseline --no-c4 --no-parity --steps 10 > gpurun_out/bench_r2w_$v.json 2>/dev/null
done
for v in good bad; do
  RVPT_B200_LIB=$V/lib$v.so timeout 200 python tools/timeline.py --batch 64 > gpurun_out/timeline_r2w_$v.md 2>&1
  echo "== $v"; grep -E "^\| (1|2|3|4|14|15) " gpurun_out/timeline_r2w_$v.md
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2w_*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f.split("r2w_")[1][:-5], "value", round(d["value"]), "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
PY
