for b in 1024 2048 4096 8192 16384; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 5 --no-c2 --no-eager-gpu --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', d['config']['batch_per_gpu'], 'samples/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'conv frac', round(d['roofline']['frac'],4), d['roofline']['forward_ms_by_kind'])"
done
