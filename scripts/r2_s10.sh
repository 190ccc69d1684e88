#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k onepass 2>&1 | tail -4
for ts in 1 0; do echo "== TSTORE=$ts"
LDCONV_OP_TSTORE=$ts timeout 300 python benchmarks/onepass_ab.py --iters 7 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'layer' in d: print(d['layer'], d.get('onepass_us'), d.get('equal'), end=' | ')
    else: print(d)
"
done
