#!/bin/bash
# round 2, GPU session b: new parity tests (fp32-class plan, N=1000, multi-step corrector, scale_by_sigma, C5, weight swap)
# then the round-1 suite on the rewritten kernels, then a short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -q -s > gpurun_out/r02b_pytest_round2.log 2>&1; echo "pytest round2 rc=$?"
grep -E 'passed|failed|FAILED|Error|error|fp32|bf16|C5|scale_by' gpurun_out/r02b_pytest_round2.log | tail -40
timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_round2.py > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E 'passed|failed|FAILED|Error|forward|sampler ' gpurun_out/r02b_pytest_gpu.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-c2 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; cat gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
