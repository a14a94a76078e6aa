#!/bin/bash
# round 2, GPU call V: non-steady interior runs out of line (A/B), parity suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -4 gpurun_out/r2v_pytest.log
{
for v in default nsc0 default; do
  L=isce3_b200/libisce3_b200_backproject.so
  [ $v != default ] && L=isce3_b200/csrc/build/variants/lib_$v.so
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 120 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2v_perf.log
