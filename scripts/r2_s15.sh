#!/bin/bash
mkdir -p gpurun_out
python benchmarks/profile_step_insitu.py --per-launch 2>/dev/null > gpurun_out/r2_step_insitu_s15.txt; head -14 gpurun_out/r2_step_insitu_s15.txt
