#!/bin/bash
# round 2, GPU call X (N GPUs): the driver's scaling command, both arms
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -10
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/r2x_bench_n$N.err; head -c 1500 gpurun_out/r2x_bench_n$N.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2x_ref_n$N.json 2> gpurun_out/r2x_ref_n$N.err
echo "reference arm N=$N rc=$?"; head -c 700 gpurun_out/r2x_ref_n$N.json; echo
