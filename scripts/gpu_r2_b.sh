#!/bin/bash
# round 2, GPU call B: ncu --set full of the steady (sub8) K=9 kernel at 1/2-scale C2
mkdir -p gpurun_out
export ISCE3_B200_LIB=isce3_b200/csrc/build/variants/lib_sub8.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_fast -s 1 -c 1 \
  -o gpurun_out/prof_fast_r2_sub8 -f python scripts/perf_fast.py 0.5 sub8 > gpurun_out/ncu_r2_sub8.log 2>&1
tail -3 gpurun_out/ncu_r2_sub8.log
