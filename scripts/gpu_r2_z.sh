#!/bin/bash
# round 2, GPU call Z: early device-to-host copies + slab ramp: suite, timeline, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
I3B_DEBUG_TIMING=1 timeout 300 python scripts/e2e_breakdown.py c2 both 3 2>&1 | tail -12 | tee gpurun_out/r2z_e2e_c2.log
I3B_DEBUG_TIMING=1 timeout 300 python scripts/e2e_breakdown.py c4 pinned 3 2>&1 | tail -5 | tee gpurun_out/r2z_e2e_c4.log
timeout 900 python bench.py --no-cpu --no-ref-cuda > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py z < gpurun_out/r2z_bench.json
