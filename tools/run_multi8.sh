#!/bin/bash
# 8-GPU check of the batch-sharded bench (torchrun, NCCL)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 3 --no-c2 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "rc=$?"; cat gpurun_out/bench_8gpu.json | cut -c1-400; tail -3 gpurun_out/bench_8gpu.err
