#!/bin/bash
# Round-2 evidence session on one B200: parity log, smoke, the bench lines, ncu launch list / full captures / DRAM traffic.
# Every ncu command runs only after the same program exited 0 without the profiler.
mkdir -p gpurun_out
T=r02
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E 'passed|failed' gpurun_out/${T}_pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 999 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference.json
# C3 guidance-weight sweep (BASELINE configs[2]): scalar w and per-sample w = 4 U[0,1]
: > gpurun_out/${T}_c3_weight_sweep.jsonl
for w in 0 0.5 1 2 4; do
  timeout 300 python bench.py --steps 20 --warmup 3 --weight $w --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e >> gpurun_out/${T}_c3_weight_sweep.jsonl 2>> gpurun_out/${T}_bench.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --weight 4 --per-sample-weight --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e >> gpurun_out/${T}_c3_weight_sweep.jsonl 2>> gpurun_out/${T}_bench.err
python -c "
import json
for ln in open('gpurun_out/${T}_c3_weight_sweep.jsonl'):
    d=json.loads(ln); print(d['config'].get('guidance','w = 1.5'), round(d['value'],1), 'samples/s')"
# ncu: launch list of the profiling target (2 PC iterations at B=8192)
timeout 200 python tools/prof_target.py > gpurun_out/${T}_prof_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/prof_target.py > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu list rc=$?"
# ncu: full sections for the conv kernel (the second forward's first 12 conv launches) and the attention block
timeout 200 python tools/prof_target.py > /dev/null 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 38 -c 8 -f -o gpurun_out/prof_conv python tools/prof_target.py > gpurun_out/${T}_ncu_conv.log 2>&1
echo "ncu conv rc=$?"; tail -2 gpurun_out/${T}_ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_block -s 5 -c 1 -f -o gpurun_out/prof_attn python tools/prof_target.py > gpurun_out/${T}_ncu_attn.log 2>&1
echo "ncu attn rc=$?"; tail -2 gpurun_out/${T}_ncu_attn.log
# ncu: DRAM bytes of every conv launch of two forwards (bench.py's roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_gemm -c 76 --csv --log-file gpurun_out/conv_traffic.csv python tools/prof_target.py > gpurun_out/${T}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
# C2: timings without a profiler, then one full capture per HBM-bound kernel
timeout 600 python tools/bench_c2.py --out gpurun_out/${T}_c2_microbench.jsonl > gpurun_out/${T}_c2.log 2>&1; echo "c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:reflect_kernel -s 3 -c 1 -o gpurun_out/prof_c2_reflect python tools/bench_c2.py --reps 1 --only reflect > gpurun_out/${T}_ncu_c2a.log 2>&1; echo "ncu reflect rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:score_hk_stream -s 3 -c 1 -o gpurun_out/prof_c2_hk_images python tools/bench_c2.py --reps 1 --only "sigma=0.1" > gpurun_out/${T}_ncu_c2b.log 2>&1; echo "ncu hk images rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:score_hk_stream -s 3 -c 1 -o gpurun_out/prof_c2_hk_mixed python tools/bench_c2.py --reps 1 --only logU > gpurun_out/${T}_ncu_c2c.log 2>&1; echo "ncu hk mixed rc=$?"
# gpurun copies back at most 64 MiB: summarise on the box (same script that writes profiles/ at home), keep only small reports
RD_SUMMARY_OUT=gpurun_out/profiles_out python tools/summarise_ncu.py ${T} gpurun_out/launches.csv gpurun_out/prof_conv.ncu-rep gpurun_out/prof_attn.ncu-rep gpurun_out/prof_c2_reflect.ncu-rep gpurun_out/prof_c2_hk_images.ncu-rep gpurun_out/prof_c2_hk_mixed.ncu-rep > gpurun_out/${T}_summarise.log 2>&1; echo "summarise rc=$?"
ls -la gpurun_out/*.ncu-rep
find gpurun_out -name "*.ncu-rep" -size +12M -delete
du -sh gpurun_out
