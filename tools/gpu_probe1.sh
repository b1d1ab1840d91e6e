#!/bin/bash
mkdir -p gpurun_out
B="--steps 2 --warmup 1 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 1"
for w in 2 8; do
echo "== MSNE_BVH=$w"
MSNE_BVH=$w python bench.py --steps 3 --warmup 3 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['avg_launch_ms'],3))"
done
MSNE_LOG_BOUNCES=1 python bench.py --steps 1 --warmup 0 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 2>&1 | grep "bounce" | head -30 > gpurun_out/bounce_counts.txt
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum
for w in 2 8; do
MSNE_BVH=$w ncu --metrics $M --clock-control none -k regex:k_trace -c 4 --csv --log-file gpurun_out/trace_first4_bvh$w.csv python bench.py --steps 1 --warmup 0 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
done
