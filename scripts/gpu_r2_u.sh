#!/bin/bash
# round 2, GPU call U: which of the run-loop changes costs time (A/B)
mkdir -p gpurun_out
{
for v in head c0n0 c1n0 c0n1 default head; do
  L=isce3_b200/libisce3_b200_backproject.so
  [ $v != default ] && L=isce3_b200/csrc/build/variants/lib_$v.so
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2u_perf.log
