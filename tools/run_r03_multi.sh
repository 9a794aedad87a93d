#!/bin/bash
# Round-2 closing multi-GPU lines (tag r03) on one 8-GPU box: C4 weak + strong scaling, C5 on 8 GPUs.  Launch exactly like the driver does.
mkdir -p gpurun_out
run() {  # run N tag args...
  N=$1; TAG=$2; shift 2
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r03_scale_${TAG}.json 2> gpurun_out/r03_scale_${TAG}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N "$@" > gpurun_out/r03_scale_${TAG}.json 2> gpurun_out/r03_scale_${TAG}.err
  fi
  echo "$TAG rc=$?"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/r03_scale_${TAG}.json').read().strip().splitlines()[-1]); print('  ', d['n_gpus'], 'GPUs', d['scaling'], d['config']['global_batch'], 'samples:', round(d['value'],1), 'samples/s', round(d['ms_per_step'],3), 'ms/step, all-gather', d['run']['all_gather_ms'], 'ms')
except Exception as e: print('  parse error', e)"
  tail -2 gpurun_out/r03_scale_${TAG}.err | cut -c1-300
}
COMMON="--steps 20 --warmup 5 --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e"
run 8 c4_weak_8gpu $COMMON                                   # 8 x 8192 = 65536 = BASELINE config C4
run 8 c5_8gpu --config c5 --steps 10 --warmup 3 --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e   # 16384 over 8 GPUs
