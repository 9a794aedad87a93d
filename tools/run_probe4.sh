#!/bin/bash
# accumulator residency: N=128 / N=64 MMA rate vs how often the issue order switches accumulator tiles (tools/probe_umma2.cu orders 10-13)
mkdir -p gpurun_out; : > gpurun_out/probe4.log
for args in "0 128 0 2 0 0 -9 148 1" "0 128 0 2 10 0 -9 148 1" "0 128 0 2 12 0 -9 148 1" "0 128 0 2 11 0 -9 148 1" "0 128 0 2 13 0 -9 148 1" \
            "0 64 0 4 12 0 -9 148 1" "0 64 0 4 13 0 -9 148 1" "0 64 0 4 4 0 -9 148 1" "0 256 0 1 12 0 -9 148 1" "0 256 0 1 0 0 -9 148 1"; do
  timeout 30 ./build/probe_umma2 $args 2>&1 | grep PROBE2 >> gpurun_out/probe4.log || echo "exit=$? args=$args" >> gpurun_out/probe4.log
done
cat gpurun_out/probe4.log
