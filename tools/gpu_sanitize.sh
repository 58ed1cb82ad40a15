#!/bin/bash
# memcheck + racecheck + synccheck on the smoke() workload (small images, all kernels)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
