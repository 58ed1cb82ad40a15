#!/bin/bash
# quick loop: parity tests + bench lines (no ncu)
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_n1.json'))
r=d['roofline']
print('default pose: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/frame', r['frame_ms_in_timed_region'], 'kernel ms', r['ms_per_launch'], 'frac', round(r['frac'],3), 'cpu', d['cpu_baseline'], 'clocks', d['clocks'])
PY
tail -3 gpurun_out/bench_${TAG}_n1.err
python bench.py --steps 30 --warmup 3 --pose pinned --no-cpu-baseline > gpurun_out/bench_${TAG}_pinned.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --scene cornell --no-cpu-baseline > gpurun_out/bench_${TAG}_cornell.json 2>/dev/null
python - <<PY
import json
for n in ('pinned','cornell'):
    d=json.load(open('gpurun_out/bench_${TAG}_%s.json'%n))
    r=d['roofline']
    print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/frame', r['frame_ms_in_timed_region'], 'frac', round(r['frac'],3), r['active_per_bounce'])
PY
