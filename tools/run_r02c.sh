#!/bin/bash
# per-launch table (bf16 + fp32-class plan) and the ncu launch list of the same target
mkdir -p gpurun_out
timeout 300 python tools/gpu_optime.py > gpurun_out/r02c_optime_bf16.log 2>&1; echo "optime rc=$?"
cat gpurun_out/r02c_optime_bf16.log
RDB200_PRECISION=fp32 RD_PROF_B=4096 timeout 300 python tools/gpu_optime.py > gpurun_out/r02c_optime_fp32.log 2>&1; echo "optime fp32 rc=$?"
head -5 gpurun_out/r02c_optime_fp32.log
