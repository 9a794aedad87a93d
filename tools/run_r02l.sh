#!/bin/bash
mkdir -p gpurun_out
for v in all none; do
  RD_FUSE_IDENTITY=$v timeout 300 python tools/gpu_optime.py > gpurun_out/r02l_optime_ident_$v.log 2>&1; echo "ident=$v rc=$?"; grep -E "B=|by kind" gpurun_out/r02l_optime_ident_$v.log
done
RD_CONV_SC_RES2=0 RD_FUSE_IDENTITY=none timeout 300 python tools/gpu_optime.py > gpurun_out/r02l_optime_nores2.log 2>&1; grep -E "B=|by kind" gpurun_out/r02l_optime_nores2.log
python - <<'PY'
import re
def load(f):
    d={}
    for ln in open(f):
        m=re.match(r"\s+(\S+)\s+([0-9.]+) ms", ln)
        if m: d[m.group(1)]=float(m.group(2))
    return d
a,b,c=load("gpurun_out/r02l_optime_ident_all.log"),load("gpurun_out/r02l_optime_ident_none.log"),load("gpurun_out/r02l_optime_nores2.log")
for k in b:
    if k in a and ("Conv_0" not in k) and "." in k and "NIN" not in k and "attn" not in k: print("%-20s ident_all %.3f  ident_none %.3f  nores2 %.3f" % (k, a[k], b[k], c.get(k, float('nan'))))
PY
