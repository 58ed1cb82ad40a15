#!/bin/bash
# instruction counts of the batched frame kernel for library variants
mkdir -p gpurun_out
for v in "$@"; do
  RVPT_B200_LIB=$PWD/rvpt_b200/variants/lib$v.so timeout 600 ncu --clock-control none -k regex:k_frame -s 3 -c 1 \
    --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio \
    --csv --log-file gpurun_out/ncu_ab_$v.csv python bench.py --no-cpu-baseline --no-c4 --no-parity --steps 1 --frames 16 > /dev/null 2>&1
  echo "== $v"; grep k_frame gpurun_out/ncu_ab_$v.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin): print(r[-3], r[-2], r[-1])"
done
