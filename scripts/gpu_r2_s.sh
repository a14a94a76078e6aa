#!/bin/bash
# round 2, GPU call S: suite, perf table, ncu --set full of the K=9 kernel, bench line, launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -4 gpurun_out/r2s_pytest.log
{
  timeout 120 python scripts/perf_fast.py 0.5 k9 2>&1 | tail -1
  I3B_FAST_NO_IMM=1 timeout 120 python scripts/perf_fast.py 0.5 k9-noimm 2>&1 | tail -1
  timeout 120 python scripts/perf_fast.py 1.0 c5k8 8 c5 2>&1 | tail -1
  timeout 120 python scripts/perf_fast.py 1.0 c5k16 16 c5 2>&1 | tail -1
  timeout 120 python scripts/perf_fast.py 1.0 c5k32 32 c5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2s_perf.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_fast -s 1 -c 1 \
  -o gpurun_out/prof_fast_r2_final -f python scripts/perf_fast.py 0.5 k9 > gpurun_out/ncu_r2_final.log 2>&1
tail -2 gpurun_out/ncu_r2_final.log
timeout 1200 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; cat gpurun_out/r2s_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2s_bench_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2s_launches.csv
