#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_s8_bench.json 2> gpurun_out/r2_s8_bench.err; tail -5 gpurun_out/r2_s8_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_s8_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')})
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','input','result')})
print('train', {k:d['config4_train'].get(k) for k in ('value','ms_per_step','error')})
print('eager', d.get('gpu_eager_baseline')); print('cpu', d.get('cpu_baseline',{}).get('value'))
r=d['roofline']; print('roof', r['achieved'], r['frac'], r['us_per_launch'], r['all_onepass_launches'], r['round1_two_kernel_path']['us_per_step'])
PY
python -m pytest tests -q -m gpu 2>&1 | tail -5
