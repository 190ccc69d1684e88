#!/bin/bash
for ts in 1 0; do echo "== TSTORE=$ts"; LDCONV_ZC_TSTORE=$ts python benchmarks/conv_ab.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l)
    print(d if 'total_us' in d else (d['cin'], d['cout'], d['hw'], d['us']), end=' | ')
print()"; done
LDCONV_ZC_TSTORE=1 python -m pytest tests/test_gpu_conv.py -q -m gpu 2>&1 | tail -2
