#!/bin/bash
mkdir -p gpurun_out
timeout 300 python benchmarks/one_kernel.py --kernel conv3x3 --cin 32 --cout 64 --hw 160 --reps 3 > gpurun_out/onek_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_zc -s 1 -c 1 -f -o gpurun_out/prof_zc32_64_r2 \
    python benchmarks/one_kernel.py --kernel conv3x3 --cin 32 --cout 64 --hw 160 --reps 3 > gpurun_out/onek_ncu.log 2>&1
echo "ncu exit $?"
