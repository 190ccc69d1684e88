#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_inference or first_layer" 2>&1 | tail -5
timeout 300 python benchmarks/l0_ab.py 2>&1 | tee gpurun_out/r2_l0_ab.jsonl
