#!/bin/bash
# round 2, GPU call L: parity suite, one-shot timeline, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -15 gpurun_out/r2l_pytest.log
I3B_DEBUG_TIMING=1 timeout 600 python scripts/e2e_breakdown.py c2 both 2 2>&1 | tee gpurun_out/r2l_e2e_c2.log
I3B_DEBUG_TIMING=1 timeout 600 python scripts/e2e_breakdown.py c4 both 2 2>&1 | tee gpurun_out/r2l_e2e_c4.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2l_bench.err
