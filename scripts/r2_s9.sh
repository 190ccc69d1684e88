#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nms.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --train-steps 0 > gpurun_out/r2_s9_bench.json 2> gpurun_out/r2_s9_bench.err; tail -3 gpurun_out/r2_s9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_s9_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step')})
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','input','result')})
PY
