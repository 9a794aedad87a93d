#!/bin/bash
# score_hk rewrite + unit entry points + codec: parity, then the C2 microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_next_rows.py -m gpu -q -s > gpurun_out/r02f_pytest_elem.log 2>&1; echo "pytest elem rc=$?"
grep -E 'passed|failed|FAILED|Error|error' gpurun_out/r02f_pytest_elem.log | tail -20
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "standalone or weight_swap" > gpurun_out/r02f_pytest_unit.log 2>&1; echo "pytest unit rc=$?"
grep -E 'passed|failed|FAILED|Error|error|resblock|attnblock' gpurun_out/r02f_pytest_unit.log | tail -20
timeout 600 python tools/bench_c2.py --out gpurun_out/r02f_c2_microbench.jsonl > gpurun_out/r02f_c2.log 2>&1; echo "c2 rc=$?"; cat gpurun_out/r02f_c2.log | tail -30
