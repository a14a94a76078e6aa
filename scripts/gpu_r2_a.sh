#!/bin/bash
# round 2, GPU call A: parity suite + kernel variant timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for v in base sub8 sub4 sub16; do
  L=isce3_b200/csrc/build/variants/lib_$v.so
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 300 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k32 32 c5 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2a_variants.log
