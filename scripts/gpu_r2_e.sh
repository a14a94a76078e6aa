#!/bin/bash
# round 2, GPU call E: full parity suite, variant timing, first run of the new bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
for v in default k32s s4; do
  L=isce3_b200/csrc/build/variants/lib_$v.so
  [ $v = default ] && L=isce3_b200/libisce3_b200_backproject.so
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 300 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 300 python scripts/perf_fast.py 1.0 $v-c5k32 32 c5 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2e_variants.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2e_bench.err
