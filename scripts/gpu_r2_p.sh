#!/bin/bash
# round 2, GPU call P: full suite + bench with the SEG=128 build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -8 gpurun_out/r2p_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2p_bench.err
