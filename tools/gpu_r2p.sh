#!/bin/bash
# descending frame groups: tests, timelines (1 GPU / one rank of eight / 4K one rank of eight), bench vs uniform groups
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2p_tests.log 2>&1
grep -E "passed|failed|error|real|differ" gpurun_out/r2p_tests.log | tail -8
timeout 200 python tools/timeline.py --batch 64 > gpurun_out/timeline_r2p_builtin_b64.md 2>&1
timeout 200 python tools/timeline.py --batch 64 --nranks 8 > gpurun_out/timeline_r2p_builtin_b64_rank0of8.md 2>&1
timeout 200 python tools/timeline.py --batch 16 --width 3840 --height 2160 --nranks 8 > gpurun_out/timeline_r2p_4k_b16_rank0of8.md 2>&1
timeout 200 python tools/timeline.py --batch 64 --pose pinned > gpurun_out/timeline_r2p_pinned_b64.md 2>&1
for f in gpurun_out/timeline_r2p_*.md; do echo "== $f"; grep -E "^\| (2|3|4|14|15) " $f; done
run() { tag=$1; shift; env "$@" timeout 600 python bench.py $ARGS --no-cpu-baseline --no-c4 > gpurun_out/bench_r2p_$tag.json 2> gpurun_out/bench_r2p_$tag.err; }
for w in builtin pinned cornell; do
  case $w in builtin) ARGS="";; pinned) ARGS="--pose pinned";; cornell) ARGS="--scene cornell --steps 5";; esac
  run ${w}_new A=1
  run ${w}_g4 RVPT_B200_FRAME_GROUP=4
  run ${w}_g16 RVPT_B200_FRAME_GROUP=16
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2p_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("r2p_")[1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d["parity_ok"], "ms/launch", round(d["roofline"]["ms_per_launch"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
