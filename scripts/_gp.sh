python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:accumulate_fast -s 1 -c 1 -o gpurun_out/prof_fast_v9 python scripts/perf_fast.py 0.5 v9 > gpurun_out/ncu_v9.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v9.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-ref-cuda > gpurun_out/bench_ncu.log 2>&1
python bench.py > gpurun_out/bench.log 2>&1
python bench.py --config c5 --taps 32 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c5k32.log 2>&1
