#!/bin/bash
# ncu full captures: k_flow and k_frame on the pinned pose and default pose
mkdir -p gpurun_out
export RVPT_B200_EXTRA_FLAGS=0x20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_flow -s 6 -c 1 \
    -f -o gpurun_out/prof_flow_pinned python bench.py --steps 1 --warmup 3 --frames 4 --pose pinned --no-cpu-baseline --graph off > gpurun_out/ncu_flow.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_flow -s 6 -c 1 \
    -f -o gpurun_out/prof_flow_default python bench.py --steps 1 --warmup 3 --frames 4 --no-cpu-baseline --graph off >> gpurun_out/ncu_flow.log 2>&1
unset RVPT_B200_EXTRA_FLAGS
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 6 -c 1 \
    -f -o gpurun_out/prof_frame_pinned python bench.py --steps 1 --warmup 3 --frames 4 --pose pinned --no-cpu-baseline --graph off > gpurun_out/ncu_frame.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 6 -c 1 \
    -f -o gpurun_out/prof_frame_default python bench.py --steps 1 --warmup 3 --frames 4 --no-cpu-baseline --graph off >> gpurun_out/ncu_frame.log 2>&1
tail -3 gpurun_out/ncu_flow.log gpurun_out/ncu_frame.log
ls -la gpurun_out/*.ncu-rep
