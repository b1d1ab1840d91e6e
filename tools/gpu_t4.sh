#!/bin/bash
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,launch__grid_size,launch__registers_per_thread
for v in park0 hyb; do
MSNE_LIB=$PWD/build/lib_$v.so ncu --metrics $M --clock-control none -k regex:k_trace -c 5 --csv --log-file gpurun_out/trace5_$v.csv python bench.py --steps 1 --warmup 0 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
done
