#!/bin/bash
# A/B of the trace kernels on cfg3: binary BVH (MSNE_BVH=2) vs compressed 8-wide BVH (default)
mkdir -p gpurun_out
for rep in 1 2; do
for w in 2 8; do
echo "== MSNE_BVH=$w"
MSNE_BVH=$w python bench.py --steps 3 --warmup 3 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | tee -a gpurun_out/ab_bvh.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['avg_launch_ms'],3), d['roofline'].get('trace_share'))"
done
done
