#!/bin/bash
# 2-GPU check of the batch-sharded bench (torchrun, NCCL)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err
echo "ref rc=$?"; cat gpurun_out/bench_2gpu_ref.json
