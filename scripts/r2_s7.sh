#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nms.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_s7_pytest.log; tail -25 gpurun_out/r2_s7_pytest.log
