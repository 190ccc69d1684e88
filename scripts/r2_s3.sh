#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k onepass 2>&1 | tail -15 > gpurun_out/r2_s3_pytest.log
tail -15 gpurun_out/r2_s3_pytest.log
timeout 300 python benchmarks/onepass_ab.py > gpurun_out/r2_s3_ab.jsonl 2> gpurun_out/r2_s3_ab.err; tail -3 gpurun_out/r2_s3_ab.err
cat gpurun_out/r2_s3_ab.jsonl
