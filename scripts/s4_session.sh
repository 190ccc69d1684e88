#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/s4_model_tests.log 2>&1; echo "model+conv tests exit $?"; tail -3 gpurun_out/s4_model_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s4e.json 2> gpurun_out/bench_s4e.err; echo "bench exit $?"; cut -c1-120 gpurun_out/bench_s4e.json
python benchmarks/profile_step_insitu.py > gpurun_out/step_insitu_s4.txt 2>&1; head -14 gpurun_out/step_insitu_s4.txt
