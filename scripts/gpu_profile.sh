#!/bin/bash
# per-kernel micro-benchmarks + ncu launch list of exactly one model step
set -u
TAG=${1:-p1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python benchmarks/ldconv_layers.py --bwd > $OUT/layers_$TAG.jsonl 2> $OUT/layers_$TAG.err
echo "layers exit $?"; tail -3 $OUT/layers_$TAG.err
timeout 300 python benchmarks/profile_step.py > $OUT/profstep_plain_$TAG.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/step_launches_$TAG.csv python benchmarks/profile_step.py > $OUT/profstep_ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/profstep_ncu_$TAG.log
