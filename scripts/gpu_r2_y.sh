#!/bin/bash
# round 2, GPU call Y (N GPUs): one-shot call timeline on every rank while all ranks upload at once
N=${1:-8}
mkdir -p gpurun_out
I3B_DEBUG_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 2 --warmup 2 --no-extras --no-cpu --no-ref-cuda > gpurun_out/r2y_bench_n$N.json 2> gpurun_out/r2y_bench_n$N.err
echo "rc=$?"; grep "i3b" gpurun_out/r2y_bench_n$N.err | tail -48; cat gpurun_out/r2y_bench_n$N.json | head -c 1200
