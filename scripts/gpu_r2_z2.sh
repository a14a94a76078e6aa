#!/bin/bash
# round 2, GPU call Z2: three-part final launch: suite, timeline, full bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z2_pytest.log
tail -4 gpurun_out/r2z2_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py z2 < gpurun_out/r2z2_bench.json
