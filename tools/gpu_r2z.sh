#!/bin/bash
# unsorted appends spread over the eight binned sub-queues (one counter per 128-byte line) + opaque scratch address:
# tests, then new vs the last committed build (variants/libgood.so) on one box, + the phase timeline
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2z_tests.log 2>&1
grep -E "passed|failed|error|real|differ" gpurun_out/r2z_tests.log | tail -8
run() { tag=$1; shift; env "$@" timeout 600 python bench.py $ARGS --no-cpu-baseline --no-c4 --no-parity > gpurun_out/bench_r2z_$tag.json 2> gpurun_out/bench_r2z_$tag.err; }
GOOD=$PWD/rvpt_b200/variants/libgood.so
for w in builtin pinned cornell; do
  case $w in builtin) ARGS="";; pinned) ARGS="--pose pinned";; cornell) ARGS="--scene cornell --steps 5";; esac
  run ${w}_new A=1
  run ${w}_good RVPT_B200_LIB=$GOOD
  run ${w}_noopq RVPT_B200_LIB=$PWD/rvpt_b200/variants/libnoopq.so
done
timeout 200 python tools/timeline.py --batch 64 > gpurun_out/timeline_r2z_builtin_b64.md 2>&1
grep -E "^\| (2|3|4|14|15) " gpurun_out/timeline_r2z_builtin_b64.md
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2z_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("r2z_")[1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
