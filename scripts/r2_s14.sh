#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --train-steps 0 > gpurun_out/r2_s14_bench.json 2> gpurun_out/r2_s14_bench.err; tail -3 gpurun_out/r2_s14_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_s14_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')})
print('e2e', d['e2e']['value'], d['e2e']['raw_head_output'])
PY
