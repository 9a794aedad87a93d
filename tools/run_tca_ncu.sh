#!/bin/bash
# ncu --set full capture of the attention-block kernel (one launch), summarised on the box (the .ncu-rep is too big to ship)
TAG=${1:-tca}
mkdir -p gpurun_out/profiles_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_block -s 5 -c 1 -f -o gpurun_out/prof_attn python tools/prof_target.py > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "ncu attn rc=$?"; tail -2 gpurun_out/${TAG}_ncu_attn.log
ncu -i gpurun_out/prof_attn.ncu-rep --page raw --csv > gpurun_out/${TAG}_attn_raw.csv 2> /dev/null
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${TAG}_attn_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "sm__pipe_tensor", "smsp__issue_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts", "smsp__pcsamp_warps_issue_stalled",
        "sm__inst_executed_pipe", "l1tex__data_bank", "dram__bytes", "lts__t_sector_hit", "sm__warps_active", "launch__", "smsp__inst_executed_op_shared", "sm__throughput",
        "l1tex__throughput", "sm__mio", "smsp__warp_issue_stalled", "l1tex__lsu_writeback", "sm__inst_executed_pipe_uniform", "idc__", "sm__pipe_shared", "l1tex__data_pipe")
with open("gpurun_out/profiles_out/${TAG}_prof_attn_ncu_summary.txt", "w") as f:
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in keep) or h == "Kernel Name":
            f.write("%-110s %-14s %s\n" % (h, u, v))
print(open("gpurun_out/profiles_out/${TAG}_prof_attn_ncu_summary.txt").read()[:200])
PY
ncu -i gpurun_out/prof_attn.ncu-rep --page source --csv > gpurun_out/profiles_out/${TAG}_attn_source.csv 2> /dev/null; ls -la gpurun_out/profiles_out/
rm -f gpurun_out/prof_attn.ncu-rep gpurun_out/${TAG}_attn_raw.csv
