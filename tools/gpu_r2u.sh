#!/bin/bash
# list loop with an opaque scratch address (no per-iteration re-derivation): tests, then new vs the previous build (one box)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2u_tests.log 2>&1
grep -E "passed|failed|error|real|differ" gpurun_out/r2u_tests.log | tail -8
run() { tag=$1; shift; env "$@" timeout 600 python bench.py $ARGS --no-cpu-baseline --no-c4 --no-parity > gpurun_out/bench_r2u_$tag.json 2> gpurun_out/bench_r2u_$tag.err; }
PREV=$PWD/rvpt_b200/variants/libprev.so
for w in builtin pinned cornell tridel; do
  case $w in builtin) ARGS="";; pinned) ARGS="--pose pinned";; cornell) ARGS="--scene cornell --steps 5";; tridel) ARGS="--scene tridel --frames 16 --steps 3";; esac
  run ${w}_new A=1
  run ${w}_prev RVPT_B200_LIB=$PREV RVPT_B200_FRAME_GROUP=16
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2u_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("r2u_")[1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
