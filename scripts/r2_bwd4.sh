#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "offset_conv_backward or accumulator or train" 2>&1 | tail -2
python benchmarks/sweep.py --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l)
    if r['dtype']=='bfloat16' and r['N']==9: print({k:r.get(k) for k in ('C','N','H','s','fwd_ms','fwd_bwd_ms')})
"
python bench.py --steps 5 --warmup 3 --train-steps 6 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d['config4_train'].get(k) for k in ('value','ms_per_step','error')})"
