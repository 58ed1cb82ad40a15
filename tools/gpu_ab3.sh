#!/bin/bash
mkdir -p gpurun_out
for v in v0 v3 v4; do
  echo "== $v"
  RVPT_B200_LIB=$PWD/rvpt_b200/variants/lib$v.so timeout 200 python tools/timeline.py --batch 16 --scene cornell 2>&1 | grep -E "^\| (2|4|6|8|14|15) "
done
