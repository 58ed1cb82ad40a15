#!/bin/bash
# phase timelines of the batched launch with leaf lists: one GPU, and one rank of eight
mkdir -p gpurun_out
timeout 200 python tools/timeline.py --batch 64 > gpurun_out/timeline_r2n_builtin_b64.md 2>&1
timeout 200 python tools/timeline.py --batch 64 --nranks 8 > gpurun_out/timeline_r2n_builtin_b64_rank0of8.md 2>&1
RVPT_B200_FRAME_GROUP=8 timeout 200 python tools/timeline.py --batch 64 --nranks 8 > gpurun_out/timeline_r2n_builtin_b64_rank0of8_g8.md 2>&1
RVPT_B200_FRAME_GROUP=32 timeout 200 python tools/timeline.py --batch 64 --nranks 8 > gpurun_out/timeline_r2n_builtin_b64_rank0of8_g32.md 2>&1
RVPT_B200_EXTRA_FLAGS=0x800 timeout 200 python tools/timeline.py --batch 64 --nranks 8 > gpurun_out/timeline_r2n_builtin_b64_rank0of8_off.md 2>&1
timeout 200 python tools/timeline.py --batch 16 --width 3840 --height 2160 --nranks 8 > gpurun_out/timeline_r2n_4k_b16_rank0of8.md 2>&1
for f in gpurun_out/timeline_r2n_*.md; do echo "== $f"; grep -E "^\| (1|2|3|4|14|15) " $f; done
