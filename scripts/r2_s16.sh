#!/bin/bash
mkdir -p gpurun_out
timeout 900 python benchmarks/config5.py > gpurun_out/r2_config5.jsonl 2> gpurun_out/r2_config5.err; tail -3 gpurun_out/r2_config5.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_config5.jsonl'):
    d=json.loads(l); print(d['p_conv_weight_sigma'], d['p_conv_bias_sigma_px'], d['ms_per_step'], d['images_per_s'], d['finite'], [(r['layer'], r['halo_miss_frac'], r['out_of_image_frac']) for r in d['ldconv_rows']])
PY
