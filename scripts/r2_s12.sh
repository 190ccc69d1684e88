#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k onepass 2>&1 | tail -3
timeout 300 python benchmarks/onepass_ab.py --iters 7 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'layer' in d: print(d['layer'], d.get('onepass_us'), d.get('equal'), end=' | ')
    else: print(d)
"
python benchmarks/trace_onepass.py --layer 15 2>&1 | sed -n 14,42p
