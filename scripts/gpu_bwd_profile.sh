#!/bin/bash
# ncu launch list of ONE LDConv training forward+backward through the module at a model layer shape
set -u
OUT=gpurun_out; mkdir -p $OUT
for L in "$@"; do
timeout 300 python benchmarks/one_bwd.py --layer $L > $OUT/onebwd_plain_L$L.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/bwd_launches_L$L.csv python benchmarks/one_bwd.py --layer $L > $OUT/onebwd_ncu_L$L.log 2>&1
echo "L$L exit $?"
done
