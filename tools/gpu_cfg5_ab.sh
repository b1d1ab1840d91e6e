#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "picker or wide" 2>&1 | tail -5
run() { echo "== $*"; local t0=$(date +%s); env "$@" python bench.py --workload cfg5 --steps 2 --warmup 1 --batch-spp 8 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],2), 'trace avg ms', round(d['roofline']['avg_launch_ms'],3), 'share', round(d['roofline']['kernel_share_of_step'],3))
    else: print(l.strip())"; echo "wall $(( $(date +%s) - t0 )) s"; }
run MSNE_BVH=2
run MSNE_BVH=8
