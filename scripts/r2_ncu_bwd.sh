#!/bin/bash
# launch list of one LDConv fwd+bwd at layer 1 (batch 64) and a --set full capture of the scatter kernel (bf16 accumulator)
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/r2_bwd_launches_L1.csv python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_ncu.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gather_bwd -c 1 -f -o $OUT/prof_scatterL1_acc16 \
    python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_ncu2.log 2>&1
echo "set-full exit $?"
