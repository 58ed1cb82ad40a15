#!/bin/bash
# memcheck + racecheck + synccheck on the smoke() workload (small images; k_frame with octant copies and
# the wave forecast), memcheck + synccheck again with the barrier-free k_flow kernel
mkdir -p gpurun_out
TAG=${1:-r01}
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitizer_$tool.log | tail -1)"
done
export RVPT_B200_EXTRA_FLAGS=0x20
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_flow_$tool.log 2>&1
  echo "flow $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitizer_flow_$tool.log | tail -1)"
done
