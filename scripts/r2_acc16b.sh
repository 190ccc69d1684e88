#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "accumulator or train_bf16 or scatter" 2>&1 | tail -3
python benchmarks/ldconv_layers.py --bwd --iters 5 2>/dev/null | grep gather_bwd > gpurun_out/r2_scatter_acc16.jsonl; grep bf16_acc gpurun_out/r2_scatter_acc16.jsonl | cut -c1-160
