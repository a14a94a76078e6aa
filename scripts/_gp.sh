python scripts/gpu_quick.py > gpurun_out/quick.log 2>&1
python scripts/perf_fast.py 0.5 v6 > gpurun_out/perf.log 2>&1
ISCE3_B200_LIB=isce3_b200/csrc/build/variants/lib_H0.so python scripts/perf_fast.py 0.5 v6_nohoist >> gpurun_out/perf.log 2>&1
I3B_FAST_NO_IMM=1 python scripts/perf_fast.py 0.5 v6_bank >> gpurun_out/perf.log 2>&1
