#!/bin/bash
# round 2, GPU call W: full suite, smoke, bench line, ncu (full set of the K=9 kernel, DRAM traffic
# of the bench launch, launch list of the bench command), sanitizer on a small scene
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
tail -4 gpurun_out/r2w_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2w_smoke.log 2>&1
tail -2 gpurun_out/r2w_smoke.log
timeout 1200 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/r2w_bench.json; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_fast -s 1 -c 1 \
  -o gpurun_out/prof_fast_r2_final -f python scripts/perf_fast.py 0.5 k9 > gpurun_out/ncu_r2_final.log 2>&1
tail -2 gpurun_out/ncu_r2_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:accumulate_fast -s 1 -c 1 \
  -o gpurun_out/prof_fast_r2_final_k16 -f python scripts/perf_fast.py 1.0 k16 16 c5 > gpurun_out/ncu_r2_final_k16.log 2>&1
tail -2 gpurun_out/ncu_r2_final_k16.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
  --clock-control none -k regex:accumulate_fast -s 3 -c 1 --csv --log-file gpurun_out/r2w_traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-ref-cuda --no-extras > gpurun_out/r2w_traffic.log 2>&1
echo "traffic rc=$?"; tail -5 gpurun_out/r2w_traffic.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-ref-cuda --no-extras > gpurun_out/r2w_bench_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2w_launches.csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitizer_scene.py > gpurun_out/r2w_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2w_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitizer_scene.py > gpurun_out/r2w_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r2w_racecheck.log
