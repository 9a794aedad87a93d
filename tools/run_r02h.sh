#!/bin/bash
# full GPU parity suite + per-launch table + short bench (after a kernel change)
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E 'passed|failed|FAILED|Error' gpurun_out/${TAG}_pytest_gpu.log | tail -8
timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_optime.log 2>&1; echo "optime rc=$?"
grep -E "B=|by kind" gpurun_out/${TAG}_optime.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print('samples/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'conv frac', round(d['roofline']['frac'],4), d['roofline']['forward_ms_by_kind'])"
