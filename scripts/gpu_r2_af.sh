#!/bin/bash
# round 2, GPU call AF: solvers after the code diet: parity, solve timing (with / without outlining
# the ellipsoid conversion and orbit interpolation as well)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_geometry.py -m gpu -x -q > gpurun_out/r2af_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2af_pytest.log
tail -3 gpurun_out/r2af_pytest.log
for v in default om; do
  L=isce3_b200/libisce3_b200_backproject.so
  [ $v != default ] && L=isce3_b200/csrc/build/variants/lib_$v.so
  for c in c4 c2; do ISCE3_B200_LIB=$L timeout 300 python scripts/e2e_breakdown.py $c pinned 1 2>&1 | grep resident | sed "s/^/$v $c /"; done
done | tee gpurun_out/r2af_solve.log
