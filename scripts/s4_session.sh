#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/s4_model_tests.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/s4_model_tests.log
LDCONV_C2F_DUAL_OUT=0 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s4_nodual.json 2> gpurun_out/bench_s4_nodual.err; echo "bench exit $?"; cut -c1-120 gpurun_out/bench_s4_nodual.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s4_dual.json 2> gpurun_out/bench_s4_dual.err; echo "bench exit $?"; cut -c1-120 gpurun_out/bench_s4_dual.json
