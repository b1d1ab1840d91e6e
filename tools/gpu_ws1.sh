#!/bin/bash
mkdir -p gpurun_out
export MSNE_BVH=2
MSNE_TRACE_WS=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_furnace.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -4
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), 'trace avg ms', round(d['roofline']['avg_launch_ms'],3), 'share', round(d['roofline']['kernel_share_of_step'],3))"; }
run MSNE_TRACE_WS=0
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=8 MSNE_LEAF_MIN=16
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=32 MSNE_LEAF_MIN=32
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=32 MSNE_LEAF_MIN=16
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=8 MSNE_LEAF_MIN=32
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=4 MSNE_LEAF_MIN=8
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=16 MSNE_LEAF_MIN=24
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=12 MSNE_LEAF_MIN=12
run MSNE_TRACE_WS=1 MSNE_REFILL_MIN=8 MSNE_LEAF_MIN=1
