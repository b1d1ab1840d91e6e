#!/bin/bash
mkdir -p gpurun_out
B=$PWD/build
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/ab_tests.log
cat gpurun_out/ab_tests.log
{
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --e2e-steps 0 --steps 3 --warmup 3 $EXTRA 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step; k_trace avg launch', round(d['roofline']['avg_launch_ms'],3), 'ms, share', round(d['roofline']['kernel_share_of_step'],3), 'reorder', round(d['roofline']['reorder_share_of_step'],3), 'launches', d['gpu_launches'])"; }
run A=default
run MSNE_LIB=$B/lib_nomc.so
run A=default2
} > gpurun_out/ab_r01h.log 2>&1
cat gpurun_out/ab_r01h.log
