#!/bin/bash
# a few throughput metrics of the batched frame kernel on one scene. usage: gpu_ncu_scene.sh <bench args...>
mkdir -p gpurun_out
timeout 900 ncu --clock-control none -k regex:k_frame -s 3 -c 1 \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio \
  --csv --log-file gpurun_out/ncu_scene.csv python bench.py --no-cpu-baseline --no-c4 --no-parity --steps 1 "$@" > /dev/null 2>&1
grep k_frame gpurun_out/ncu_scene.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin): print(r[-3], r[-2], r[-1])"
