#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/probe2.log
for args in "0 64 0 3 4 0 -9 1 0" "0 64 0 3 4 0 -9 1 1" "0 64 0 3 4 0 -9 148 0" "0 64 0 3 4 0 -9 148 1" "0 128 0 2 0 0 0 148 1" "0 256 0 1 0 0 0 148 1"; do
  timeout 30 ./build/probe_umma2 $args 2>&1 | grep PROBE2 >> gpurun_out/probe2.log || echo "exit=$? args=$args" >> gpurun_out/probe2.log
done
cat gpurun_out/probe2.log
