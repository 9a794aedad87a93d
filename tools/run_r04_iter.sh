#!/bin/bash
# round-2 late iteration loop: network parity tests, per-launch table with and without the switch under test
TAG=${1:-r04}
SWITCH=${2:-RD_CONV_POLY}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_round2.py -m gpu -q -x -k "not N1000" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_optime.log 2>&1; echo "optime rc=$?"
grep -E "B=|by kind|forward" gpurun_out/${TAG}_optime.log
env ${SWITCH}=0 timeout 300 python tools/gpu_optime.py > gpurun_out/${TAG}_optime_off.log 2>&1; echo "optime (${SWITCH}=0) rc=$?"
grep -E "B=|by kind|forward" gpurun_out/${TAG}_optime_off.log
