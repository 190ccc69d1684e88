#!/bin/bash
for d in 0 1 2 3 4 8 5 7 15; do
  echo -n "dbg=$d: "
  LDCONV_OP_DBG=$d timeout 300 python benchmarks/onepass_ab.py --iters 5 --layers 1,15,10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'layer' in d: print(d['layer'], d.get('onepass_us'), end=' | ')
print()
"
done
