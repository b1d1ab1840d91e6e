#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/ab_tests.log
( MSNE_SORT_THRESHOLD=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) >> gpurun_out/ab_tests.log
cat gpurun_out/ab_tests.log
{
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --e2e-steps 0 $EXTRA 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step; k_trace avg launch', round(d['roofline']['avg_launch_ms'],3), 'ms, share', round(d['roofline']['kernel_share_of_step'],3), 'reorder', round(d['roofline']['reorder_share_of_step'],3), 'launches', d['gpu_launches'])"; }
EXTRA="--steps 4 --warmup 3" run A=cfg3
EXTRA="--steps 4 --warmup 3" run MSNE_SORT_THRESHOLD=0
EXTRA="--workload cfg5 --batch-spp 8 --steps 4 --warmup 3" run A=cfg5
} > gpurun_out/ab_r01h.log 2>&1
cat gpurun_out/ab_r01h.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_run.log 2>&1
