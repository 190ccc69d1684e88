#!/bin/bash
# first-layer rows kernel: parity tests, A/B timing, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_inference or first_layer_rows" > gpurun_out/l0_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/l0_tests.log
LDCONV_L0_CONST=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_inference or first_layer_rows" > gpurun_out/l0_tests_smem.log 2>&1; echo "tests (shared-memory variant) exit $?"; tail -1 gpurun_out/l0_tests_smem.log
for cfg in "0 6 1" "1 6 0" "1 6 1" "1 5 1"; do set -- $cfg; LDCONV_L0_ROWS=$1 LDCONV_L0_MINB=$2 LDCONV_L0_CONST=$3 python benchmarks/l0_ab.py >> gpurun_out/l0_ab.jsonl 2>gpurun_out/l0_ab.err; done
cat gpurun_out/l0_ab.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_l0.json 2> gpurun_out/bench_l0.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_l0.json
