#!/bin/bash
# C5 bench (bf16 + fp32 plan), fp32-plan C3 bench
mkdir -p gpurun_out
timeout 900 python bench.py --config c5 --steps 10 --warmup 3 --no-eager-gpu > gpurun_out/r02e_bench_c5.json 2> gpurun_out/r02e_bench_c5.err; echo "c5 rc=$?"; cat gpurun_out/r02e_bench_c5.json; tail -3 gpurun_out/r02e_bench_c5.err
timeout 900 python bench.py --precision fp32 --steps 20 --warmup 3 --no-eager-gpu --no-c2 --no-cpu-baseline > gpurun_out/r02e_bench_c3_fp32.json 2> gpurun_out/r02e_bench_c3_fp32.err; echo "c3 fp32 rc=$?"; cat gpurun_out/r02e_bench_c3_fp32.json; tail -3 gpurun_out/r02e_bench_c3_fp32.err
RD_PROF_CFG=c5 timeout 600 python tools/gpu_optime.py > gpurun_out/r02e_optime_c5.log 2>&1; echo "optime c5 rc=$?"; head -80 gpurun_out/r02e_optime_c5.log
