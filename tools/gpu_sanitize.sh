#!/bin/bash
# memcheck + racecheck + synccheck on the smoke() workload (small images; k_frame with octant copies and
# the wave forecast) and on a batched render_frames launch
mkdir -p gpurun_out
TAG=${1:-r01}
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitizer_$tool.log | tail -1)"
done
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/batch_smoke.py > gpurun_out/${TAG}_sanitizer_batch_$tool.log 2>&1
  echo "batch $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitizer_batch_$tool.log | tail -1)"
done
