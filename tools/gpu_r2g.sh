#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 6 -c 1 \
    -f -o gpurun_out/prof_frame_cornell python bench.py --steps 1 --warmup 3 --frames 4 --scene cornell --no-cpu-baseline --graph off > gpurun_out/ncu_cornell.log 2>&1
tail -n 3 gpurun_out/ncu_cornell.log
