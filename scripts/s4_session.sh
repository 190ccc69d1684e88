#!/bin/bash
mkdir -p gpurun_out
for c in 64 48 64 48; do echo "pack max C = $c"; LDCONV_GEMM_PACK_MAX_C=$c python bench.py --steps 30 --warmup 5 2>/dev/null | cut -c1-90; done
python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -1
