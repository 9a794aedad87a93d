#!/bin/bash
# C2 microbench session: timings (no profiler), then ncu captures of the HBM-bound kernels
mkdir -p gpurun_out
timeout 600 python tools/bench_c2.py --out gpurun_out/c2.jsonl > gpurun_out/c2.log 2>&1; echo "c2 rc=$?"; cat gpurun_out/c2.log | cut -c1-260
if [ "${DO_NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:'reflect_kernel|score_hk_kernel|pc_predictor|pc_corrector|pc_norms' -s 3 -c 1 -o gpurun_out/prof_c2_reflect python tools/bench_c2.py --reps 1 --only reflect > gpurun_out/ncu_c2a.log 2>&1; echo "ncu reflect rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:'score_hk_kernel' -s 3 -c 1 -o gpurun_out/prof_c2_hk python tools/bench_c2.py --reps 1 --only logU > gpurun_out/ncu_c2b.log 2>&1; echo "ncu hk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:'pc_predictor' -s 3 -c 1 -o gpurun_out/prof_c2_pred python tools/bench_c2.py --reps 1 --only predictor > gpurun_out/ncu_c2c.log 2>&1; echo "ncu pred rc=$?"
fi
