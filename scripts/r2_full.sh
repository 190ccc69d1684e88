#!/bin/bash
# whole GPU suite + the bench line (the round-end driver does the same on a fresh box)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_full_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_full_bench.json 2> gpurun_out/r2_full_bench.err; tail -3 gpurun_out/r2_full_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_full_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step','lib_sha16')})
print('e2e', d['e2e']['value'], d['e2e'].get('raw_head_output'))
print('train', {k:d.get('config4_train',{}).get(k) for k in ('value','ms_per_step','error')})
r=d['roofline']; print('roof', r['achieved'], r['frac'], r['us_per_launch'], r['all_onepass_launches'], r['round1_two_kernel_path']['us_per_step'])
PY
