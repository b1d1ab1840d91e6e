#!/bin/bash
mkdir -p gpurun_out
export MSNE_BVH=2
for s in 0 1; do
MSNE_SORT=$s MSNE_SORT_OBITS=5 MSNE_SORT_DBITS=5 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_sort$s.csv python bench.py --steps 1 --warmup 0 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
done
MSNE_TRACE_WS=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k full_size 2>&1 | tail -30 > gpurun_out/ws_fail.txt
