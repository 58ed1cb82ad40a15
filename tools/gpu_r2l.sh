#!/bin/bash
# leaf-list primary wave (primary_phase_beam): GPU tests, then A/B on ONE box of the frame-group size and of
# the flag that turns the lists off (RVPT_B200_FLAG_NO_LEAF_LISTS = 0x800).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2l_tests.log 2>&1
grep -E "passed|failed|error|real|differ" gpurun_out/r2l_tests.log | tail -8
run() { # tag, env..., -- bench args
  tag=$1; shift
  env "$@" timeout 300 python bench.py $ARGS --no-cpu-baseline --no-c4 > gpurun_out/bench_r2l_$tag.json 2> gpurun_out/bench_r2l_$tag.err
}
for w in builtin pinned cornell; do
  case $w in builtin) ARGS="";; pinned) ARGS="--pose pinned";; cornell) ARGS="--scene cornell --steps 5";; esac
  run ${w}_off RVPT_B200_EXTRA_FLAGS=0x800
  for g in 4 8 16 32 64; do run ${w}_g$g RVPT_B200_FRAME_GROUP=$g; done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2l_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("r2l_")[1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
