#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_s4_pytest.log; tail -3 gpurun_out/r2_s4_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_s4_bench.json 2> gpurun_out/r2_s4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_s4_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step','gpu_eager_baseline')}); print(d['e2e']['value'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldconv_onepass -s 1 -c 1 -f -o gpurun_out/prof_onepassL1_v1 \
    python benchmarks/onepass_ab.py --layers 1 --iters 2 > gpurun_out/ncu_onepassL1_v1.log 2>&1
echo "ncu exit $?"
