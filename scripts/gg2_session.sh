#!/bin/bash
# One gpurun call for the gather+GEMM kernel: parity tests, per-layer A/B over the kernel's environment switches, then the bench.
# LDCONV_GG_V=1 first-generation kernel / 2 specialised v2 (default) / 3 v2 on run-time geometry; LDCONV_GG_MERGE=0 two-barrier
# flow; LDCONV_GG_TG=2|3 threads per CTA / 128; LDCONV_GG_PLAN, LDCONV_GG_CTAS buffering plan / CTAs per SM.
set -u
TAG=${1:-gg2}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest (gather+GEMM, module, model) ==="
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -m gpu -q -x --timeout 300 > $OUT/pytest_$TAG.log 2>&1
echo "exit $?"; tail -3 $OUT/pytest_$TAG.log
LDCONV_GG_V=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "gather_gemm or far_outside" > $OUT/pytest_v3_$TAG.log 2>&1
echo "run-time geometry instances: exit $?"; tail -2 $OUT/pytest_v3_$TAG.log
run() { echo "$*"; env "$@" timeout 300 python benchmarks/ldconv_layers.py 2>&1 | grep gather_gemm | python -c "
import sys,json
print(' '.join('L%d:%s' % (d['layer'], d['us']) for d in map(json.loads,sys.stdin)))"; }
run LDCONV_GG_V=2
run LDCONV_GG_V=2 LDCONV_GG_MERGE=0
run LDCONV_GG_V=1
echo "=== bench ==="
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "exit $?"; cut -c1-200 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
