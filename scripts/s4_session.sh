#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/s4_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s4_tests.log
LDCONV_SPPF_CASCADE=0 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "sppf" 2>&1 | tail -1
LDCONV_DECODE_STAGED=0 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "decode" 2>&1 | tail -1
LDCONV_SPPF_CASCADE=0 LDCONV_DECODE_STAGED=0 python benchmarks/profile_step_insitu.py 2>/dev/null | grep -E "launches|decode|sppf"
python benchmarks/profile_step_insitu.py 2>/dev/null | grep -E "launches|decode|sppf"
LDCONV_SPPF_CASCADE=0 LDCONV_DECODE_STAGED=0 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-90
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s4g.json 2> gpurun_out/bench_s4g.err; cut -c1-90 gpurun_out/bench_s4g.json
