#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_next_rows.py -m gpu -q -s -k "ode" > gpurun_out/r02j_pytest_ode.log 2>&1; echo "pytest ode rc=$?"
grep -E 'passed|failed|FAILED|Error|ode |floor' gpurun_out/r02j_pytest_ode.log | tail -12
timeout 600 python tools/bench_c2.py --only "score_hk" --out gpurun_out/r02j_c2.jsonl 2>&1 | python -c "
import sys,re
for ln in sys.stdin:
    m=re.search(r'\"case\": \"([^\"]+)\".*?\"ms_avg\": ([0-9.]+).*?\"frac\": ([0-9.]+)', ln)
    if m: print('%-48s %.4f ms  frac %.3f' % (m.group(1), float(m.group(2)), float(m.group(3))))
"
