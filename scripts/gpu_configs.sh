#!/bin/bash
# BASELINE.json configs 2, 4, 5 on one B200: layer sweep (fwd / fwd+bwd / GEMM TFLOP/s), training step, 1280x1280 layers
set -u
TAG=${1:-c1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python benchmarks/sweep.py > $OUT/sweep_$TAG.jsonl 2> $OUT/sweep_$TAG.err; echo "sweep exit $?"; tail -2 $OUT/sweep_$TAG.err
timeout 600 python benchmarks/train_step.py --batch 128 --steps 10 > $OUT/train_$TAG.json 2> $OUT/train_$TAG.err; echo "train exit $?"; tail -3 $OUT/train_$TAG.err; cat $OUT/train_$TAG.json
timeout 600 python benchmarks/ldconv_layers.py --scale 2 --batch 32 --bwd > $OUT/layers1280_$TAG.jsonl 2> $OUT/layers1280_$TAG.err; echo "1280 exit $?"; tail -2 $OUT/layers1280_$TAG.err
