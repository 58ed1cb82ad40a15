#!/bin/bash
# ncu full capture of the slow build under bench.py (to compare with prof_frame_r02 = the good build)
mkdir -p gpurun_out
V=$PWD/rvpt_b200/variants
RVPT_B200_LIB=$V/libbad.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 4 -c 1 -f -o gpurun_out/prof_frame_bad \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-parity > gpurun_out/ncu_frame_bad.log 2>&1
tail -2 gpurun_out/ncu_frame_bad.log
