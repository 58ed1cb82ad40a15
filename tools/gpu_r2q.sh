#!/bin/bash
# leaf lists from the launch's first phase vs built per unit, by frame-group size (one box)
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py $ARGS --no-cpu-baseline --no-c4 --no-parity > gpurun_out/bench_r2q_$tag.json 2> gpurun_out/bench_r2q_$tag.err; }
for w in builtin pinned; do
  case $w in builtin) ARGS="";; pinned) ARGS="--pose pinned";; esac
  for g in 4 8 16 32; do
    run ${w}_pre_g$g RVPT_B200_FRAME_GROUP=$g
    run ${w}_inl_g$g RVPT_B200_FRAME_GROUP=$g RVPT_B200_LIST_INLINE=1
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2q_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("r2q_")[1][:-5], "value", round(d["value"]), "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
