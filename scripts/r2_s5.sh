#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 -1" "2 0" "2 1" "3 -1" "3 1" "3 3"; do set -- $cfg
  echo "== TG=$1 plan=$2"
  LDCONV_OP_TG=$1 LDCONV_OP_PLAN=$2 timeout 300 python benchmarks/onepass_ab.py --iters 5 --layers 1,15,10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'layer' in d: print(d['layer'], d.get('onepass_us'), d.get('equal'), end=' | ')
    else: print(d)
"
done
