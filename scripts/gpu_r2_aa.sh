#!/bin/bash
# round 2, GPU call AA: phase-rate guard in the kernel: parity + kernel timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2aa_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
tail -3 gpurun_out/r2aa_pytest.log
{
  timeout 120 python scripts/perf_fast.py 0.5 k9 2>&1 | tail -1
  timeout 120 python scripts/perf_fast.py 1.0 c5k8 8 c5 2>&1 | tail -1
  timeout 120 python scripts/perf_fast.py 1.0 c5k16 16 c5 2>&1 | tail -1
} | tee gpurun_out/r2aa_perf.log
