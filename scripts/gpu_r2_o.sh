#!/bin/bash
mkdir -p gpurun_out
for v in default seg64; do
  L=isce3_b200/csrc/build/variants/lib_$v.so
  [ $v = default ] && L=isce3_b200/libisce3_b200_backproject.so
  ISCE3_B200_LIB=$L timeout 300 python scripts/parity_quick.py $v
done 2>&1 | tee gpurun_out/r2o_parity.log
