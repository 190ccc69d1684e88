#!/bin/bash
# gather v2 + 32-bit scatter backward: tests, per-layer timings (with the generic gather as A/B), one-step ncu launch list
set -u
TAG=${1:-s3b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log
timeout 600 python benchmarks/ldconv_layers.py --bwd > $OUT/layers_$TAG.jsonl 2> $OUT/layers_$TAG.err
echo "layers exit $?"
python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT/layers_$TAG.jsonl") if l.startswith("{")]
for k in ("gather_fwd","gather_bwd","gather_gemm_fwd"):
    print(k, " ".join("L%d:%s(%s)" % (r["layer"], r["us"], int(r["GBps"])) for r in rows if r["kernel"]==k and r.get("variant") in ("tma_tile","tcgen05",None,"atomics","")))
PY
LDCONV_GATHER_V=1 timeout 300 python benchmarks/ldconv_layers.py 2>&1 | grep '"gather_fwd", "variant": "tma_tile"' | python -c "
import sys,json
print('generic gather', ' '.join('L%d:%s' % (d['layer'], d['us']) for d in map(json.loads,sys.stdin)))"
timeout 300 python benchmarks/profile_step.py > $OUT/profstep_plain_$TAG.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/step_launches_$TAG.csv python benchmarks/profile_step.py > $OUT/profstep_ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/profstep_ncu_$TAG.log
