#!/bin/bash
# round 2, session 1: new parity tests + whole GPU suite + bench baseline
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2_s1_pytest.log
tail -5 gpurun_out/r2_s1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s1_bench.json 2> gpurun_out/r2_s1_bench.err
cut -c1-400 gpurun_out/r2_s1_bench.json
