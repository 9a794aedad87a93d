#!/bin/bash
# quick loop for the tcgen05 attention block: network parity tests, per-phase cycle profile, per-launch table
TAG=${1:-tca}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_round2.py -m gpu -q -x -k "not c5 and not N1000" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/${TAG}_pytest.log
bash tools/run_tcaprof.sh $TAG
timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_optime.log 2>&1; echo "optime rc=$?"
grep -E "B=|by kind|attn" gpurun_out/${TAG}_optime.log
