#!/bin/bash
# experiment: which stage bounds the tcgen05 GEMM (1 = no operand TMA traffic, 2 = no epilogue math / stores)
for d in 0 1 2 3; do echo "LDCONV_GEMM_DBG=$d"; LDCONV_GEMM_DBG=$d timeout 300 python benchmarks/ldconv_layers.py 2>&1 | grep '"gemm_fwd", "variant": "tcgen05"' | python -c "
import sys,json
print(' '.join('L%d:%s' % (d['layer'], d['us']) for d in map(json.loads,sys.stdin)))"; done
