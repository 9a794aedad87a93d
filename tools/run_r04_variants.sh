#!/bin/bash
# parity tests on the product library, then the per-launch table for the product and for measurement builds (librdb200_<v>.so)
TAG=${1:-r04}; shift
mkdir -p gpurun_out
L=$PWD/optimized-diffusion-model_b200/rdb200
timeout 1200 python -m pytest tests/test_gpu_network.py tests/test_gpu_round2.py -m gpu -q -x -k "not N1000" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_optime.log 2>&1; echo "product rc=$?"
grep -E "B=|by kind|forward|${SHOW:-up_blocks.6.Conv_0|out_head|_attn.0}" gpurun_out/${TAG}_optime.log
for v in "$@"; do
  RDB200_LIB=$L/librdb200_$v.so timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_optime_$v.log 2>&1; echo "$v rc=$?"
  grep -E "B=|by kind|${SHOW:-up_blocks.6.Conv_0|out_head|_attn.0}" gpurun_out/${TAG}_optime_$v.log
done
