#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python tools/gpu_optime.py > gpurun_out/optime_$name.log 2>&1; echo "== $name: $(sed -n 1,3p gpurun_out/optime_$name.log | tr '\n' ' ')"; }
run base RD_X=0
run s20 "RD_CONV_FORCE_S=128,128,2,20;256,128,2,10"
run s20b "RD_CONV_FORCE_S=128,128,2,20"
run s14 "RD_CONV_FORCE_S=128,128,2,14;256,128,2,10"
