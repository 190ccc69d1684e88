#!/bin/bash
# experiment: which stage bounds the gather+GEMM kernel (results are wrong on purpose when a stage is skipped)
for d in 0 1 2 3 4 8 7 15; do
echo "dbg $d (1=no phase2, 2=no phase1, 4=no TMA, 8=no epilogue math/stores)"; LDCONV_GG_DBG=$d timeout 300 python benchmarks/ldconv_layers.py 2>&1 | grep gather_gemm | python -c "
import sys,json
print(' '.join('L%d:%s' % (d['layer'], d['us']) for d in map(json.loads,sys.stdin)))"
done
