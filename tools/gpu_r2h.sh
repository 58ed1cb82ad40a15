#!/bin/bash
mkdir -p gpurun_out
for t in 2 4 8 16; do
  export RVPT_B200_TAIL_RAYS_PER_WARP=$t
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t${t}_n1.json 2>/dev/null
  timeout 300 python bench.py --steps 30 --warmup 3 --pose pinned --no-cpu-baseline > gpurun_out/bench_t${t}_pinned.json 2>/dev/null
  timeout 300 python bench.py --steps 10 --warmup 3 --scene cornell --bounces 16 --no-cpu-baseline > gpurun_out/bench_t${t}_cornell.json 2>/dev/null
done
python - <<PY
import json
for v in (2,4,8,16):
  for n in ('n1','pinned','cornell'):
    try:
        d=json.load(open('gpurun_out/bench_t%d_%s.json'%(v,n)))
        r=d['roofline']
        print('tail', v, n, 'value', round(d['value']), 'us/frame', round(r['frame_ms_in_timed_region']*1000,1), r['active_per_bounce'][:5])
    except Exception as e:
        print(v, n, 'failed', e)
PY
