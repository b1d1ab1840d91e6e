#!/bin/bash
# A/B helper used for the sweeps in profiles/r01_sweeps.md: each argument is an "ENV=val[,ENV=val...]" set, run on the
# short bench configuration (cfg3, batch 8 spp, 16 Mi-path waves). Example (on the GPU box):
#   bash tools/ab.sh A=0 MSNE_MAX_LEAF=2 MSNE_LIB=$PWD/build/lib_variant.so
for cfg in "$@"; do
  echo "== $cfg"
  env ${cfg//,/ } python bench.py --steps 3 --warmup 2 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step; k_trace avg launch', round(d['roofline']['avg_launch_ms'],3), 'ms, share', round(d['roofline']['kernel_share_of_step'],3))"
done
