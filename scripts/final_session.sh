#!/bin/bash
# One gpurun call for the consolidated evidence of a session: full GPU tests, smoke, both bench arms, the ncu launch list of
# the bench command, one `ncu --set full` capture of the dominant LDConv kernel, per-layer kernel timings.
# Usage (repo root, GPU box): bash scripts/final_session.sh [tag]
set -u
TAG=${1:-fin}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu ==="
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_$TAG.log 2>&1
echo "exit $?"; tail -4 $OUT/pytest_$TAG.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "exit $?"; tail -3 $OUT/smoke_$TAG.log
echo "=== bench ==="
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
BRC=$?
echo "exit $BRC"; cut -c1-700 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "=== bench --impl reference ==="
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo "exit $?"; cut -c1-600 $OUT/bench_ref_$TAG.json
echo "=== per-layer kernels ==="
timeout 600 python benchmarks/ldconv_layers.py --bwd > $OUT/layers_$TAG.jsonl 2> $OUT/layers_$TAG.err
echo "exit $?"
if [ $BRC -eq 0 ]; then
  echo "=== ncu launch list of exactly one step (cudaProfilerStart/Stop around it; the same script run plain first) ==="
  # `ncu -c N python bench.py` only catches the model-initialisation kernels (several hundred element-wise launches)
  timeout 300 python benchmarks/profile_step.py > $OUT/profstep_plain_$TAG.log 2>&1 &&
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file $OUT/step_launches_$TAG.csv python benchmarks/profile_step.py > $OUT/profstep_ncu_$TAG.log 2>&1
  echo "exit $?"; tail -2 $OUT/profstep_ncu_$TAG.log
fi
echo "=== ncu --set full: gather+GEMM kernel at layer 1, stand-alone gather at layer 1 ==="
bash scripts/gpu_ncu_kernel.sh ggL1_$TAG ldconv_gg2 --kernel gg --layer 1
bash scripts/gpu_ncu_kernel.sh gatherL1_$TAG gather_fwd_tiled --kernel gather --layer 1
