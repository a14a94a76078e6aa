bash scripts/run_variants.sh 0.5 S64U2 S128U2 S64U4 > gpurun_out/variants.log 2>&1
ISCE3_B200_LIB=isce3_b200/csrc/build/variants/lib_S64U2.so python scripts/gpu_quick.py > gpurun_out/quick_S64U2.log 2>&1
ISCE3_B200_LIB=isce3_b200/csrc/build/variants/lib_S128U2.so python scripts/gpu_quick.py > gpurun_out/quick_S128U2.log 2>&1
