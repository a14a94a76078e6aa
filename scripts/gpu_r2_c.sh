#!/bin/bash
# round 2, GPU call C: kernel variant timing batch 2 + reproducibility tests with the default build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reproducible or slab or plan_resident or sharding_over_device or c2_like or airborne or every_inst or general_output or swath_edges" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
for v in s8 s8q s8qh s8qhp s8qhpw s8qhpl s4qhp; do
  L=isce3_b200/csrc/build/variants/lib_$v.so
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 300 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k32 32 c5 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2c_variants.log
