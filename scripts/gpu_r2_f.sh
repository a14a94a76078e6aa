#!/bin/bash
# round 2, GPU call F: new parity tests + one-shot call anatomy (C2, C4)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -15 gpurun_out/r2f_pytest.log
I3B_DEBUG_TIMING=1 timeout 600 python scripts/e2e_breakdown.py c2 both 3 2>&1 | tee gpurun_out/r2f_e2e_c2.log
I3B_DEBUG_TIMING=1 timeout 600 python scripts/e2e_breakdown.py c4 both 3 2>&1 | tee gpurun_out/r2f_e2e_c4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_c4.csv python scripts/e2e_breakdown.py c4 pageable 2 > /dev/null 2>&1
grep -c . gpurun_out/r2f_launches_c4.csv
