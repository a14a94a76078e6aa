#!/bin/bash
# round 2, GPU call AE: solvers with outlined samplers / single-call-site root finder: parity, solve timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_geometry.py -m gpu -x -q > gpurun_out/r2ae_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ae_pytest.log
tail -3 gpurun_out/r2ae_pytest.log
for c in c4 c2 c1; do timeout 300 python scripts/e2e_breakdown.py $c pinned 1 2>&1 | grep resident | sed "s/^/$c /"; done | tee gpurun_out/r2ae_solve.log
I3B_NO_TIGHT_DEM=1 timeout 300 python scripts/e2e_breakdown.py c4 pinned 1 2>&1 | grep resident | sed "s/^/c4 full-interval /" | tee -a gpurun_out/r2ae_solve.log
