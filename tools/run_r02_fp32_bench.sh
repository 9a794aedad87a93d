mkdir -p gpurun_out
timeout 900 python bench.py --precision fp32 --steps 20 --warmup 3 --no-eager-gpu --no-c2 --no-cpu-baseline > gpurun_out/r02_bench_c3_fp32plan.json 2> gpurun_out/r02_bench_c3_fp32plan.err; echo "fp32 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_c3_fp32plan.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['forward_ms_by_kind'])"
timeout 900 python bench.py --config c5 --precision fp32 --steps 5 --warmup 3 --no-eager-gpu --no-c2 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5_fp32plan.json 2> gpurun_out/r02_bench_c5_fp32plan.err; echo "c5 fp32 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_c5_fp32plan.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['forward_ms_by_kind'])"
