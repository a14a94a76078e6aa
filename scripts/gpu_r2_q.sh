#!/bin/bash
# round 2, GPU call Q: reduced run polynomial: accuracy (point-target and noise scenes) and speed, A/B
mkdir -p gpurun_out
for v in default prev; do
  L=isce3_b200/csrc/build/variants/lib_$v.so
  [ $v = default ] && L=isce3_b200/libisce3_b200_backproject.so
  ISCE3_B200_LIB=$L timeout 300 python scripts/parity_quick.py $v
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 120 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k32 32 c5 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2q_ab.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -6 gpurun_out/r2q_pytest.log
