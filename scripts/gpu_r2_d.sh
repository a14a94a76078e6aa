#!/bin/bash
# round 2, GPU call D: ncu --set full of the K=16 kernel on the airborne scene
mkdir -p gpurun_out
export ISCE3_B200_LIB=isce3_b200/csrc/build/variants/lib_s8q.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_fast -s 1 -c 1 \
  -o gpurun_out/prof_fast_r2_k16 -f python scripts/perf_fast.py 1.0 k16 16 c5 > gpurun_out/ncu_r2_k16.log 2>&1
tail -3 gpurun_out/ncu_r2_k16.log
