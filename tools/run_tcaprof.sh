#!/bin/bash
# Per-phase cycle counts of the tcgen05 attention block (build with -DRD_TCA_PROF into librdb200_tcaprof.so first):
#   python -c "from rdb200 import build; build.build(force=True, defines=('RD_TCA_PROF',), out=build.OUT.replace('.so','_tcaprof.so'))"
TAG=${1:-tca}
mkdir -p gpurun_out
L=$PWD/optimized-diffusion-model_b200/rdb200
RDB200_LIB=$L/librdb200_tcaprof.so timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_tcaprof.log 2>&1; echo "prof rc=$?"
grep "attn_tc" gpurun_out/${TAG}_tcaprof.log | sort | uniq -c | sort -rn | head -8
grep -E "attn|by kind" gpurun_out/${TAG}_tcaprof.log | grep -v attn_tc
