#!/bin/bash
# explicit reconvergence at the top of the frame loop of primary_phase_beam: new vs the previous build, by group size (one box)
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py $ARGS --no-cpu-baseline --no-c4 --no-parity > gpurun_out/bench_r2v_$tag.json 2> gpurun_out/bench_r2v_$tag.err; }
PREV=$PWD/rvpt_b200/variants/libprev.so
for w in builtin pinned cornell; do
  case $w in builtin) ARGS="";; pinned) ARGS="--pose pinned";; cornell) ARGS="--scene cornell --steps 5";; esac
  run ${w}_new A=1
  run ${w}_new_g8 RVPT_B200_FRAME_GROUP=8
  run ${w}_new_g32 RVPT_B200_FRAME_GROUP=32
  run ${w}_prev RVPT_B200_LIB=$PREV RVPT_B200_FRAME_GROUP=16
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2v_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("r2v_")[1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
