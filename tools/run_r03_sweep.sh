#!/bin/bash
# C3 guidance-weight sweep (BASELINE configs[2]) on the closing kernels: scalar w and per-sample w = 4 U[0,1]
mkdir -p gpurun_out
T=r03
: > gpurun_out/${T}_c3_weight_sweep.jsonl
for w in 0 0.5 1 2 4; do
  timeout 300 python bench.py --steps 20 --warmup 3 --weight $w --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e >> gpurun_out/${T}_c3_weight_sweep.jsonl 2>> gpurun_out/${T}_sweep.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --weight 4 --per-sample-weight --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e >> gpurun_out/${T}_c3_weight_sweep.jsonl 2>> gpurun_out/${T}_sweep.err
python -c "
import json
for ln in open('gpurun_out/${T}_c3_weight_sweep.jsonl'):
    d=json.loads(ln); print(d['config'].get('guidance','w = 1.5'), round(d['value'],1), 'samples/s', d['clocks'].get('sm_mhz'))"
timeout 600 python tools/bench_c2.py --out gpurun_out/${T}_c2_microbench.jsonl > gpurun_out/${T}_c2.log 2>&1; echo "c2 rc=$?"; tail -12 gpurun_out/${T}_c2.log
