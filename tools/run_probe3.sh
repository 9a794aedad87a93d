#!/bin/bash
# MMA rate (N=64, 3 tiles, 9 taps) next to a full epilogue on four other warps, and the epilogue's own cost per
# 32-column block with (reps 64) and without (order 9 = no MMAs) the tensor core running; see tools/probe_umma2.cu
mkdir -p gpurun_out; : > gpurun_out/probe3.log
for args in "4 705" "4 755" "9 755" "4 790"; do
  set -- $args
  timeout 60 ./build/probe_umma2 0 64 0 3 $1 $2 -9 148 1 2>&1 | grep PROBE2 >> gpurun_out/probe3.log || echo "exit=$? args=$args" >> gpurun_out/probe3.log
done
cat gpurun_out/probe3.log
