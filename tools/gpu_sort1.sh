#!/bin/bash
mkdir -p gpurun_out
export MSNE_BVH=2
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), 'trace avg ms', round(d['roofline']['avg_launch_ms'],3), 'share', round(d['roofline']['kernel_share_of_step'],3))"; }
run MSNE_SORT=0
run MSNE_SORT=1 MSNE_SORT_OBITS=5 MSNE_SORT_DBITS=5
run MSNE_SORT=1 MSNE_SORT_OBITS=4 MSNE_SORT_DBITS=4
run MSNE_SORT=1 MSNE_SORT_OBITS=6 MSNE_SORT_DBITS=3
run MSNE_SORT=1 MSNE_SORT_OBITS=5 MSNE_SORT_DBITS=5 MSNE_SORT_DIRFIRST=1
run MSNE_SORT=1 MSNE_SORT_OBITS=7 MSNE_SORT_DBITS=5
run MSNE_SORT=1 MSNE_SORT_OBITS=6 MSNE_SORT_DBITS=0
