#!/bin/bash
# one GPU session: parity tests, smoke, bench, launch list, one full ncu capture of the conv kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E 'passed|failed|FAILED|Error|forward|sampler ' gpurun_out/pytest_gpu.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-60} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
if [ "${DO_NCU:-1}" = "1" ]; then
timeout 200 python tools/prof_target.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/prof_target.py > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 200 python tools/prof_target.py > gpurun_out/prof_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 20 -c 4 -o gpurun_out/prof_conv python tools/prof_target.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
# DRAM bytes of every conv launch of two forwards (bench.py's roofline.traffic comes from this capture)
timeout 200 python tools/prof_target.py > gpurun_out/prof_plain3.log 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_gemm -c 96 --csv --log-file gpurun_out/conv_traffic.csv python tools/prof_target.py > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
fi
# BASELINE config C2: HBM-bound kernels (timings without a profiler, then one full capture per kernel)
bash tools/run_c2.sh
