#!/bin/bash
# ncu --set full on one kernel of the path.  usage: bash scripts/gpu_ncu_kernel.sh <tag> <regex> <one_kernel.py args...>
set -u
TAG=$1; REGEX=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python benchmarks/one_kernel.py "$@" > $OUT/onek_plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 1 -c 1 -f -o $OUT/prof_$TAG \
    python benchmarks/one_kernel.py "$@" > $OUT/onek_ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/onek_ncu_$TAG.log
