#!/bin/bash
# One gpurun call for the gather+GEMM kernel: parity tests, per-layer A/B (LDCONV_GG_V=1 generic kernel, =2 specialised,
# =3 v2 kernel without template specialisation), then the bench.  Usage: bash scripts/gg2_session.sh [tag]
set -u
TAG=${1:-gg2}
OUT=gpurun_out
mkdir -p $OUT
echo "=== pytest -m gpu ==="
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_$TAG.log 2>&1
echo "exit $?"; tail -5 $OUT/pytest_$TAG.log
run() { echo "$*"; env "$@" timeout 300 python benchmarks/ldconv_layers.py 2>&1 | grep gather_gemm | python -c "
import sys,json
print(' '.join('L%d:%s' % (d['layer'], d['us']) for d in map(json.loads,sys.stdin)))"; }
run LDCONV_GG_V=1
run LDCONV_GG_V=2
echo "=== bench ==="
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "exit $?"; tail -c 2500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
