#!/bin/bash
# round 2, GPU call R: full suite (sinc, rangecomp scaling, runtime segment), accuracy, perf, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -8 gpurun_out/r2r_pytest.log
timeout 300 python scripts/parity_quick.py default 2>&1 | tee gpurun_out/r2r_parity.log
for v in default; do
  L=isce3_b200/libisce3_b200_backproject.so
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 0.5 $v-k9 2>&1 | tail -1
  ISCE3_B200_LIB=$L I3B_FAST_NO_IMM=1 timeout 120 python scripts/perf_fast.py 0.5 $v-k9-noimm 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k8 8 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k16 16 c5 2>&1 | tail -1
  ISCE3_B200_LIB=$L timeout 120 python scripts/perf_fast.py 1.0 $v-c5k32 32 c5 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2r_perf.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitizer_scene.py > gpurun_out/r2r_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2r_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitizer_scene.py > gpurun_out/r2r_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2r_racecheck.log
