#!/bin/bash
# A/B harness for the measurement switches of DESIGN.md section 5: one per-launch table (tools/gpu_optime.py) per
# configuration, all in the same GPU session so that box-to-box noise (about +-0.5 %) does not enter the comparison.
#   gpurun -- 'bash tools/run_exp.sh'          then compare gpurun_out/optime_<name>.log column by column
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python tools/gpu_optime.py > gpurun_out/optime_$name.log 2>&1; echo "== $name: $(sed -n 1,3p gpurun_out/optime_$name.log | tr '\n' ' ')"; }
run base RD_X=0
run nopdl RD_CONV_PDL=0
run nommas RD_CONV_DEBUG=4          # timing only: skips every MMA
run nostream RD_CONV_DEBUG=64       # timing only: no filter stream
# examples of the geometry overrides measured (and rejected) in round 1:
#   run nt4   RD_CONV_SINGLE_PEN=1.0 RD_CONV_TIE_EPS=0.02
#   run w8    RD_CONV_WSTAGES=8 RD_CONV_ASTAGES_STREAM=2
#   run s20   "RD_CONV_FORCE_S=128,128,2,20;256,128,2,10"
#   run f1    RD_CONV_RES2=1 "RD_CONV_FORCE_NT=128,64,8,2"
