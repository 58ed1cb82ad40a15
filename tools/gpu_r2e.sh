#!/bin/bash
# sharded dynamic bounce waves: all tests, A/B bench (new default vs 0x40 = old static dealing), timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|words differ|Error" | tail -12
for v in ord ref; do
  case $v in
    ord) unset RVPT_B200_EXTRA_FLAGS;;
    ref) export RVPT_B200_EXTRA_FLAGS=0x80;;
  esac
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${v}_n1.json 2> gpurun_out/bench_${v}_n1.err
  timeout 300 python bench.py --steps 30 --warmup 3 --pose pinned --no-cpu-baseline > gpurun_out/bench_${v}_pinned.json 2>/dev/null
  timeout 300 python bench.py --steps 10 --warmup 3 --scene cornell --no-cpu-baseline > gpurun_out/bench_${v}_cornell.json 2>/dev/null
done
unset RVPT_B200_EXTRA_FLAGS
python - <<PY
import json
for v in ("ord","ref"):
  for n in ('n1','pinned','cornell'):
    try:
        d=json.load(open('gpurun_out/bench_%s_%s.json'%(v,n)))
        r=d['roofline']
        print(v, n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/frame', round(r['frame_ms_in_timed_region'],4), 'frac', round(r['frac'],3), r['active_per_bounce'][:4], d['clocks']['sm_mhz'])
    except Exception as e:
        print(v, n, 'failed', e)
PY
timeout 120 python tools/timeline.py > gpurun_out/timeline_builtin.md 2>&1; tail -8 gpurun_out/timeline_builtin.md
timeout 120 python tools/timeline.py --scene cornell > gpurun_out/timeline_cornell.md 2>&1; tail -17 gpurun_out/timeline_cornell.md
