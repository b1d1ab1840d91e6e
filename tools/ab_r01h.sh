#!/bin/bash
mkdir -p gpurun_out
{
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --e2e-steps 0 --steps 3 --warmup 3 $EXTRA 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:60], round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step; share', round(d['roofline']['kernel_share_of_step'],3), 'reorder', round(d['roofline']['reorder_share_of_step'],3))"; }
EXTRA="--detail 0.05 --batch-spp 32" run A=d005
EXTRA="--detail 0.05 --batch-spp 32" run MSNE_SORT_THRESHOLD=0
EXTRA="--detail 0.005 --batch-spp 32" run A=d0005
EXTRA="--detail 0.005 --batch-spp 32" run MSNE_SORT_THRESHOLD=0
EXTRA="--detail 0.0005 --batch-spp 32" run A=d00005
EXTRA="--detail 0.0005 --batch-spp 32" run MSNE_SORT_THRESHOLD=0
} > gpurun_out/ab_r01h.log 2>&1
cat gpurun_out/ab_r01h.log
