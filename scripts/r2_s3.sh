#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k onepass 2>&1 | tail -5 > gpurun_out/r2_s3_pytest.log
tail -3 gpurun_out/r2_s3_pytest.log
timeout 300 python benchmarks/onepass_ab.py > gpurun_out/r2_s3_ab.jsonl 2> gpurun_out/r2_s3_ab.err; tail -3 gpurun_out/r2_s3_ab.err
python -c "
import sys, json
for l in open('gpurun_out/r2_s3_ab.jsonl'):
    d = json.loads(l)
    if 'layer' in d: print(d['layer'], d['two_kernel_us'], d.get('onepass_us'), d.get('onepass_GBps'), d.get('equal'))
    else: print(d)
"
