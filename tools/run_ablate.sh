#!/bin/bash
# Role-cost ablations of conv_gemm_kernel: per-launch table of the product build and of the measurement builds that drop
# one role's work each (see the RD_ABL_* macros in csrc/conv_gemm.cu).  Usage on the GPU box: bash tools/run_ablate.sh TAG
TAG=${1:-abl}
mkdir -p gpurun_out
L=$PWD/optimized-diffusion-model_b200/rdb200
timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_product.log 2>&1; echo "product rc=$?"
for v in ${ABL_VARIANTS:-NO_MMA NO_EPI NO_XFORM NO_WSTREAM}; do
  RDB200_LIB=$L/librdb200_abl_$v.so timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_$v.log 2>&1; echo "$v rc=$?"
done
python - <<PY
import re
rows = {}
import os
names = ["product"] + os.environ.get("ABL_VARIANTS", "NO_MMA NO_EPI NO_XFORM NO_WSTREAM").split()
order = []
for v in names:
    for ln in open("gpurun_out/${TAG}_%s.log" % v):
        m = re.match(r"\s+(\S+)\s+([0-9.]+) ms", ln)
        if m:
            rows.setdefault(m.group(1), {})[v] = float(m.group(2))
            if v == "product": order.append((m.group(1), ln.split("ms", 1)[1].strip()[:110]))
print("%-24s " % "op" + " ".join("%8s" % n[-8:] for n in names))
tot = {v: 0.0 for v in names}
for n, info in order:
    r = rows[n]
    if "cin" not in info: continue
    print("%-24s " % n + " ".join("%8.3f" % r.get(v, float("nan")) for v in names) + "  " + info[:70])
    for v in names: tot[v] += r.get(v, 0.0)
print("%-24s " % "conv total" + " ".join("%8.3f" % tot[v] for v in names))
PY
