python scripts/gpu_quick.py > gpurun_out/quick.log 2>&1
for k in 8 16 32; do python scripts/perf_fast.py 0.5 c5k$k $k c5; done > gpurun_out/perf.log 2>&1
python scripts/perf_fast.py 0.5 c2k9 >> gpurun_out/perf.log 2>&1
