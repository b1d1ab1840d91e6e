#!/bin/bash
# Session-3 A/B batch 3 (cfg3, batch 8 spp, 16 Mi-path waves): emitter records, shade block size
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/ab_tests.log
B=$PWD/build
bash tools/ab.sh A=default MSNE_LIB=$B/lib_noem.so MSNE_LIB=$B/lib_st64.so MSNE_LIB=$B/lib_st256.so A=default2 MSNE_LIB=$B/lib_noem.so > gpurun_out/ab_r01h.log 2>&1
cat gpurun_out/ab_tests.log gpurun_out/ab_r01h.log
