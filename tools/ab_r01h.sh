#!/bin/bash
# Session-3 A/B batch 2 (cfg3, batch 8 spp, 16 Mi-path waves) + ncu source-level capture of k_shade<PBR> for stall attribution
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/ab_tests.log
B=$PWD/build
bash tools/ab.sh A=default MSNE_LIB=$B/lib_pos.so MSNE_LIB=$B/lib_minb5.so MSNE_LIB=$B/lib_minb7.so MSNE_LIB=$B/lib_minb8.so \
     MSNE_LIB=$B/lib_hld.so MSNE_LIB=$B/lib_hst.so A=default2 > gpurun_out/ab_r01h.log 2>&1
cat gpurun_out/ab_tests.log gpurun_out/ab_r01h.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 5 -c 1 -f -o gpurun_out/shade_pbr_full \
    python bench.py --steps 1 --warmup 1 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
ls -la gpurun_out
