#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -q -s > gpurun_out/r02o_pytest_round2.log 2>&1; echo "pytest round2 rc=$?"
grep -E 'passed|failed|FAILED|Error|error|fp32|bf16|C5|scale_by|resblock|attnblock' gpurun_out/r02o_pytest_round2.log | tail -40
RDB200_PRECISION=fp32 RD_PROF_B=4096 timeout 300 python tools/gpu_optime.py > gpurun_out/r02o_optime_fp32.log 2>&1; grep -E "B=|by kind" gpurun_out/r02o_optime_fp32.log
RD_PROF_CFG=c5 timeout 600 python tools/gpu_optime.py > gpurun_out/r02o_optime_c5.log 2>&1; grep -E "B=|by kind" gpurun_out/r02o_optime_c5.log
