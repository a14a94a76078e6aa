#!/bin/bash
# round 2, GPU call N (2 GPUs): bench.py under torchrun at N=2 (all arms), 2-GPU tests, non-uniform PRF test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or current_device or non_uniform or blocks_api or reproducible" > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -8 gpurun_out/r2n_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2n_bench_n2.err
