#!/bin/bash
# round 2, GPU call AC: narrow look bracket for raster DEMs: parity (all raster / geometry tests), C4 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_geometry.py -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ac_pytest.log
tail -3 gpurun_out/r2ac_pytest.log
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "c4" 2>&1 | tail -2
I3B_DEBUG_TIMING=1 timeout 300 python scripts/e2e_breakdown.py c4 pinned 3 2>&1 | tail -4 | tee gpurun_out/r2ac_e2e_c4.log
I3B_NO_TIGHT_DEM=1 timeout 300 python scripts/e2e_breakdown.py c4 pinned 2 2>&1 | grep resident | tee -a gpurun_out/r2ac_e2e_c4.log
