#!/bin/bash
# Round-2 closing evidence session on one B200 (tag r03): parity log, smoke, the bench lines, ncu launch list / full
# captures (conv, tcgen05 attention block, mma output head) / DRAM traffic of the conv launches.  Every ncu command runs
# only after the same program exited 0 without the profiler.
mkdir -p gpurun_out gpurun_out/profiles_out
T=r03
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E 'passed|failed' gpurun_out/${T}_pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/${T}_bench_reference.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_driver_flags.json 2>> gpurun_out/${T}_bench.err; echo "bench (driver flags) rc=$?"; cut -c1-400 gpurun_out/${T}_bench_driver_flags.json
timeout 900 python bench.py --steps 999 --warmup 5 --no-c2 --no-eager-gpu --no-cpu-baseline > gpurun_out/${T}_bench.json 2>> gpurun_out/${T}_bench.err; echo "bench (full pass) rc=$?"; cut -c1-400 gpurun_out/${T}_bench.json
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_c5.json 2>> gpurun_out/${T}_bench.err; echo "bench c5 rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c5.json
timeout 600 python bench.py --precision fp32 --steps 10 --warmup 3 --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_c3_fp32plan.json 2>> gpurun_out/${T}_bench.err; echo "bench fp32 plan rc=$?"; cut -c1-300 gpurun_out/${T}_bench_c3_fp32plan.json
timeout 300 python tools/gpu_optime.py > gpurun_out/${T}_optime.log 2>&1; echo "optime rc=$?"
RD_PROF_HW=9x9 timeout 300 python tools/gpu_optime.py > gpurun_out/${T}_optime_9x9.log 2>&1; echo "optime 9x9 rc=$?"
# ncu: launch list of the profiling target (2 PC iterations at B=8192)
timeout 200 python tools/prof_target.py > gpurun_out/${T}_prof_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/prof_target.py > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu list rc=$?"
# ncu: full sections for the conv kernel (first eight conv launches of the second forward), the attention block, the output head
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 38 -c 8 -f -o gpurun_out/prof_conv python tools/prof_target.py > gpurun_out/${T}_ncu_conv.log 2>&1
echo "ncu conv rc=$?"; tail -2 gpurun_out/${T}_ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_block -s 5 -c 1 -f -o gpurun_out/prof_attn python tools/prof_target.py > gpurun_out/${T}_ncu_attn.log 2>&1
echo "ncu attn rc=$?"; tail -2 gpurun_out/${T}_ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:out_head -s 1 -c 1 -f -o gpurun_out/prof_outhead python tools/prof_target.py > gpurun_out/${T}_ncu_outhead.log 2>&1
echo "ncu out_head rc=$?"; tail -2 gpurun_out/${T}_ncu_outhead.log
# ncu: DRAM bytes of every conv launch of two forwards (bench.py's roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_gemm -c 76 --csv --log-file gpurun_out/conv_traffic.csv python tools/prof_target.py > gpurun_out/${T}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
RD_SUMMARY_OUT=gpurun_out/profiles_out python tools/summarise_ncu.py ${T} gpurun_out/launches.csv gpurun_out/prof_conv.ncu-rep gpurun_out/prof_attn.ncu-rep gpurun_out/prof_outhead.ncu-rep > gpurun_out/${T}_summarise.log 2>&1; echo "summarise rc=$?"
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
