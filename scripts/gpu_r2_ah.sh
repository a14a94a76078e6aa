#!/bin/bash
# round 2, GPU call AH: early result copies for pageable arrays too: parity, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2ah_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ah_pytest.log
tail -3 gpurun_out/r2ah_pytest.log
I3B_DEBUG_TIMING=1 I3B_POOL_KEEP_MB=-1 timeout 300 python scripts/e2e_breakdown.py c2 both 3 2>&1 | grep -v "run:" | tail -8 | tee gpurun_out/r2ah_e2e_c2.log
