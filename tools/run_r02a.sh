#!/bin/bash
# round 2, GPU session a: parity suite on the shipped code, split-bf16 accumulation probe, error budget of the
# approximate SFU functions (product build vs -DRD_EXACT_ACT build of the same sources)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_probe_split.log 2>&1
for args in "576 64 1" "576 64 3" "576 64 4" "2304 128 1" "2304 128 3" "2304 128 4" "1152 128 3 2" "1728 64 3 3"; do
  timeout 60 ./build/probe_split $args >> gpurun_out/r02a_probe_split.log 2>&1
done
cat gpurun_out/r02a_probe_split.log
timeout 900 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E 'passed|failed|FAILED|Error|forward|sampler ' gpurun_out/r02a_pytest_gpu.log | tail -30
timeout 300 python tests/diag/gpu_diag.py > gpurun_out/r02a_diag_product.log 2>&1; echo "diag rc=$?"
RDB200_LIB=$PWD/optimized-diffusion-model_b200/rdb200/librdb200_exact.so timeout 300 python tests/diag/gpu_diag.py > gpurun_out/r02a_diag_exact.log 2>&1; echo "diag exact rc=$?"
grep -E "forward|tap" gpurun_out/r02a_diag_product.log | head -40
echo ---- exact
grep -E "forward|tap" gpurun_out/r02a_diag_exact.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
RDB200_LIB=$PWD/optimized-diffusion-model_b200/rdb200/librdb200_exact.so timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
