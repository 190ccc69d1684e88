#!/bin/bash
# bf16-accumulator backward: parity tests, accumulator error over offset scales, per-layer scatter timing (both accumulators), training step
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "accumulator or train_bf16 or scatter or offset_conv_backward" 2>&1 | tail -5
python benchmarks/acc16_error.py 2>&1 | tee gpurun_out/r2_acc16_error.jsonl
python benchmarks/ldconv_layers.py --bwd --iters 5 2>/dev/null | grep gather_bwd > gpurun_out/r2_scatter_acc16.jsonl; grep bf16_acc gpurun_out/r2_scatter_acc16.jsonl
python bench.py --steps 5 --warmup 3 --train-steps 6 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d['config4_train'].get(k) for k in ('value','ms_per_step','final_loss','error')})"
