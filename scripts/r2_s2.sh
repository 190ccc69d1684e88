#!/bin/bash
# round 2, session 2: whole GPU suite (no -x) + bench with the new keys + strong-scaling line at N=1
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_s2_pytest.log
tail -5 gpurun_out/r2_s2_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s2_bench.json 2> gpurun_out/r2_s2_bench.err
tail -3 gpurun_out/r2_s2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_s2_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_eager_baseline','cpu_baseline','lib_sha16')})
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['roofline']['all_gg_launches'])
PY
