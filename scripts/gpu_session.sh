#!/bin/bash
# One gpurun call: parity tests (CUDA-core GEMMs first, then with tcgen05), smoke, bench, ncu launch list.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_session.sh [tag]
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1

echo "=== pytest -m gpu, LDCONV_FORCE_FFMA=1 ==="
LDCONV_FORCE_FFMA=1 timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_ffma_$TAG.log 2>&1
echo "exit $?"; tail -5 $OUT/pytest_ffma_$TAG.log

echo "=== tcgen05 GEMM tests alone (a trap here must not poison the other tests) ==="
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm_bf16" --timeout 300 > $OUT/pytest_umma_$TAG.log 2>&1
echo "exit $?"; tail -25 $OUT/pytest_umma_$TAG.log

echo "=== pytest -m gpu (tcgen05 enabled) ==="
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_$TAG.log 2>&1
echo "exit $?"; tail -15 $OUT/pytest_$TAG.log

echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "exit $?"; tail -3 $OUT/smoke_$TAG.log

echo "=== bench (FFMA GEMM) ==="
LDCONV_FORCE_FFMA=1 timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_ffma_$TAG.json 2> $OUT/bench_ffma_$TAG.err
echo "exit $?"; tail -c 1500 $OUT/bench_ffma_$TAG.json; tail -3 $OUT/bench_ffma_$TAG.err

echo "=== bench ==="
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
BRC=$?
echo "exit $BRC"; tail -c 3000 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err

echo "=== bench --impl reference ==="
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo "exit $?"; tail -c 600 $OUT/bench_ref_$TAG.json

if [ $BRC -eq 0 ]; then
  echo "=== ncu launch list (same command run plain first) ==="
  timeout 600 python bench.py --steps 2 --warmup 3 > $OUT/plain_$TAG.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 > $OUT/ncu_launch_$TAG.log 2>&1
  echo "exit $?"; tail -2 $OUT/ncu_launch_$TAG.log
fi
