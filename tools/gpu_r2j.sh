#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
bash tools/gpu_scale.sh r2j 1
timeout 900 python tools/bounce_sweep.py r02 > gpurun_out/bounce_sweep_r02.log 2>&1; cp profiles/r02_bounce_sweep.md gpurun_out/ 2>/dev/null; head -14 profiles/r02_bounce_sweep.md | tail -6
