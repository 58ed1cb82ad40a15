#!/bin/bash
# source-level ncu capture (instruction counts per SASS line) of the batched frame kernel on one scene.
# usage: gpu_ncu_src.sh <tag> <bench args...>; read here with tools/ncu_blocks.py
mkdir -p gpurun_out
TAG=$1; shift
timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on \
  -k regex:k_frame -s 3 -c 1 -f -o gpurun_out/src_${TAG} \
  python bench.py --no-cpu-baseline --no-c4 --no-parity --steps 1 "$@" > gpurun_out/ncu_src_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_src_${TAG}.log
ls -la gpurun_out/src_${TAG}.ncu-rep
