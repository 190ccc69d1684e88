#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/s4_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s4_tests.log
for cfg in "0 256" "1 128" "1 256"; do set -- $cfg
  echo "pack=$1 max_k=$2"; LDCONV_GEMM_PACK=$1 LDCONV_GEMM_PACK_MAX_K=$2 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-90
done
python benchmarks/profile_step_insitu.py 2>/dev/null | grep -E "launches|umma_gemm"
LDCONV_GEMM_PACK=0 python benchmarks/profile_step_insitu.py 2>/dev/null | grep -E "launches|umma_gemm"
