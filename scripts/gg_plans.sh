#!/bin/bash
# experiment: gather+GEMM kernel under different CTAs/SM (and optionally buffering plans / stage skips); prints us per layer
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gather_gemm" 2>&1 | tail -3
run() { echo "$*"; env "$@" timeout 300 python benchmarks/ldconv_layers.py 2>&1 | grep gather_gemm | python -c "
import sys,json
print(' '.join('L%d:%s' % (d['layer'], d['us']) for d in map(json.loads,sys.stdin)))"; }
run LDCONV_GG_CTAS=4
run LDCONV_GG_CTAS=3
run LDCONV_GG_CTAS=2
run LDCONV_GG_MINB3=1 LDCONV_GG_CTAS=4
run LDCONV_GG_PLAN=0
run LDCONV_GG_PLAN=1
run LDCONV_GG_DBG=15
run LDCONV_GG_DBG=7
run LDCONV_GG_DBG=1
