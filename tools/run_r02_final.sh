#!/bin/bash
# Final verification of the shipped tree on a fresh box, in the order the driver uses: GPU parity suite, smoke, the two bench
# arms with the driver's flags; then the DRAM-traffic capture of the final conv sources (bench.py's roofline.traffic).
mkdir -p gpurun_out
T=r02final
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E 'passed|failed' gpurun_out/${T}_pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench.err; echo "bench ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2>> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/${T}_bench_reference.json').read().strip().splitlines()[-1])
print('ours', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], '| reference arm', round(r['value'],3), r['cpu_baseline']['kind'], '| eager gpu', round(d['eager_gpu_baseline']['value'],2), '| same config', d['config']==r['config'])"
timeout 200 python tools/prof_target.py > /dev/null 2>&1 && \
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_gemm -c 76 --csv --log-file gpurun_out/conv_traffic.csv python tools/prof_target.py > gpurun_out/${T}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/prof_target.py > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu list rc=$?"
RD_SUMMARY_OUT=gpurun_out/profiles_out python tools/summarise_ncu.py r02 gpurun_out/launches.csv /nonexistent > gpurun_out/${T}_summarise.log 2>&1; echo "summarise rc=$?"
du -sh gpurun_out
