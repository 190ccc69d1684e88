#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s4_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s4_tests.log
echo "stage off"; LDCONV_GEMM_STAGE=0 python benchmarks/gemm_pack_ab.py 2>&1 | grep '"P": 1'
echo "stage on"; python benchmarks/gemm_pack_ab.py 2>&1 | grep '"P": 1'
LDCONV_GEMM_STAGE=0 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-90
python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-90
