#!/bin/bash
# final single-GPU evidence for the current build: whole GPU suite, bench line, ncu launch list + set-full capture, A/B of the LDConv rows
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_final_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err
python benchmarks/onepass_ab.py --iters 7 > gpurun_out/r2_final_onepass_ab.jsonl 2>/dev/null; tail -1 gpurun_out/r2_final_onepass_ab.jsonl
python benchmarks/profile_step_insitu.py --per-launch 2>/dev/null > gpurun_out/r2_final_step_insitu.txt; head -12 gpurun_out/r2_final_step_insitu.txt
bash scripts/r2_capture.sh
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step','lib_sha16')})
print('e2e', d['e2e']['value'], d['e2e'].get('raw_head_output'))
print('train', {k:d.get('config4_train',{}).get(k) for k in ('value','ms_per_step','error')})
print('eager', d.get('gpu_eager_baseline')); print('cpu', d.get('cpu_baseline',{}).get('value'))
r=d['roofline']; print('roof', r['achieved'], r['frac'], r['us_per_launch'], r['all_onepass_launches'], r['round1_two_kernel_path']['us_per_step'])
PY
