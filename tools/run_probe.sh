#!/bin/bash
# runs the UMMA convention probe over the variants the conv kernel depends on
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.log 2>&1
for args in "288 64 0 0" "288 64 0 1" "289 64 0 1" "289 64 1 1" "301 64 11 1" "301 64 21 1" "288 128 0 1" "289 128 7 1" "289 192 3 1" "289 256 3 1" "289 16 3 1"; do
  timeout 30 ./build/probe_umma $args >> gpurun_out/probe.log 2>&1
  echo "exit=$? args=$args" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
