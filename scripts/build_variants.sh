#!/bin/bash
# Developer tool: build fast-kernel variants for A/B timing.  Usage: build_variants.sh name "-DX=1 ..." ...
set -e
cd "$(dirname "$0")/../isce3_b200/csrc"
make -j4 >/dev/null
mkdir -p build/variants
NV="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -ccbin /usr/bin/g++"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc $NV $flags -c accumulate_fast.cu -o build/variants/fast_$name.o
  /usr/local/cuda/bin/nvcc -shared -o build/variants/lib_$name.so build/capi.o build/solve_kernels.o build/peaks.o build/rangecomp.o build/variants/fast_$name.o -lpthread -lcufft
  echo built $name
done
