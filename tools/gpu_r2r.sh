#!/bin/bash
# run-to-run spread of the default bench line (separate processes, one box): 6 x frame groups of 16, 6 x of 8
mkdir -p gpurun_out
for g in 16 8; do for i in 1 2 3 4 5 6; do
  RVPT_B200_FRAME_GROUP=$g timeout 300 python bench.py --no-cpu-baseline --no-c4 --no-parity --steps 10 > gpurun_out/bench_r2r_g${g}_$i.json 2>/dev/null
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2r_*.json")):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f.split("r2r_")[1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
PY
