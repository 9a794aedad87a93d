#!/bin/bash
# ncu --set full on selected conv_gemm launches of the profiling target
mkdir -p gpurun_out
timeout 200 python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-2} -o gpurun_out/prof_conv_v2 -f python tools/prof_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
