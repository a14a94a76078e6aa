#!/bin/bash
# round 2, GPU call K: non-uniform PRF test, one-shot timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -15 gpurun_out/r2k_pytest.log
I3B_DEBUG_TIMING=1 timeout 600 python scripts/e2e_breakdown.py c2 both 3 2>&1 | tee gpurun_out/r2k_e2e_c2.log
I3B_DEBUG_TIMING=1 timeout 600 python scripts/e2e_breakdown.py c4 both 3 2>&1 | tee gpurun_out/r2k_e2e_c4.log
