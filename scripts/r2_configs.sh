#!/bin/bash
# BASELINE.json configs 1 and 2 on the current build: B=1 latency, the single-layer sweep, tensor-pipe capture of the K=2304 GEMM
OUT=gpurun_out; mkdir -p $OUT
python benchmarks/config1.py > $OUT/r2_config1.jsonl 2>$OUT/r2_config1.err; cat $OUT/r2_config1.jsonl
python benchmarks/sweep.py > $OUT/r2_sweep_config2.jsonl 2>$OUT/r2_sweep.err; wc -l $OUT/r2_sweep_config2.jsonl; tail -3 $OUT/r2_sweep_config2.jsonl | cut -c1-220
python benchmarks/gemm_big.py > $OUT/r2_gemm_big.jsonl 2>/dev/null; cat $OUT/r2_gemm_big.jsonl
timeout 600 ncu --set full --clock-control none -k regex:umma_gemm -s 3 -c 1 -f -o $OUT/prof_gemm_k2304 python benchmarks/gemm_big.py --iters 2 > $OUT/gemm_ncu.log 2>&1
echo "gemm set-full exit $?"
