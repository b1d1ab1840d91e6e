#!/bin/bash
# Builds moonshine_b200/libmoonshine.so (the reference's artifact name, build.zig: libmoonshine.so) for sm_100a.
#  -fmad=false            : device FP32 is evaluated exactly as written (DESIGN.md "Arithmetic")
#  -ffp-contract=off      : same for the host FP32 that feeds the device (camera basis, inverses, alias table)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${OUT:-../libmoonshine.so}
$NVCC -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false \
      -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function \
      -shared kernels.cu runtime.cpp bvh_build.cpp -ldl -o $OUT "$@"
echo "built $(realpath $OUT)"
# the native `offline` front end (offline/main.zig:27-203 over the C ABI): only next to the default library
if [ "$OUT" = "../libmoonshine.so" ] && [ $# -eq 0 ]; then
  ${CXX:-g++} -std=c++17 -O2 -Wall -Wextra -ffp-contract=off offline_main.cpp -o ../offline -L.. -lmoonshine -lz -Wl,-rpath,'$ORIGIN'
  echo "built $(realpath ../offline)"
fi
