#!/bin/bash
# round 2, final GPU call: what the driver runs (suite, smoke, bench line)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
tail -3 gpurun_out/r2final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err
echo "bench rc=$?"; python scripts/bench_summary.py final < gpurun_out/r2final_bench.json; wc -l gpurun_out/r2final_bench.json
