#!/bin/bash
# round 2, GPU call AB: steady runs chained over the 16-pulse tile (A/B), parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2ab_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ab_pytest.log
tail -3 gpurun_out/r2ab_pytest.log
timeout 300 python scripts/parity_quick.py default 2>&1 | tee gpurun_out/r2ab_parity.log
{
for v in default pair0; do
  L=isce3_b200/libisce3_b200_backproject.so
  [ $v != default ] && L=isce3_b200/csrc/build/variants/lib_$v.so
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 120 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k32 32 c5 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2ab_perf.log
