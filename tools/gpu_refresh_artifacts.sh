#!/bin/bash
# Refresh of every per-round artefact under gpurun_out/ (then: python tools/summarize_profile.py r01; cp gpurun_out/r01_timeline_*.md profiles/):
# tests + bench + ncu launch list + ncu full capture + sanitizer (gpu_profile.sh), bounce sweep, the other workloads, timelines.
bash tools/gpu_profile.sh r01
timeout 600 python tools/bounce_sweep.py r01 > gpurun_out/sweep.log 2>&1; cp profiles/r01_bounce_sweep.md gpurun_out/ 2>/dev/null; tail -2 gpurun_out/sweep.log
timeout 300 python bench.py --steps 10 --warmup 3 --scene cornell --no-cpu-baseline > gpurun_out/bench_r01_cornell.json 2>/dev/null
for n in 20000 500000; do timeout 300 python bench.py --steps 5 --warmup 3 --scene mesh --mesh-tris $n --no-cpu-baseline > gpurun_out/bench_r01_mesh$n.json 2>/dev/null; done
python - <<PY
import json
for n in ('n1','pinned','cornell','mesh20000','mesh500000'):
    try:
        d=json.load(open('gpurun_out/bench_r01_%s.json'%n)); r=d['roofline']
        print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'us/frame', round(r['frame_ms_in_timed_region']*1000,1), 'frac', round(r['frac'],3), 'Mrays/s', round(d['mrays_per_s']), r['active_per_bounce'][:4])
    except Exception as e: print(n, 'failed', e)
PY
timeout 120 python tools/timeline.py > gpurun_out/r01_timeline_builtin.md 2>&1
timeout 120 python tools/timeline.py --pose pinned > gpurun_out/r01_timeline_pinned.md 2>&1
timeout 120 python tools/timeline.py --scene cornell > gpurun_out/r01_timeline_cornell.md 2>&1
RVPT_B200_EXTRA_FLAGS=0x40 timeout 120 python tools/timeline.py > gpurun_out/r01_timeline_builtin_noforecast.md 2>&1
tail -9 gpurun_out/r01_timeline_pinned.md
