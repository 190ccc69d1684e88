#!/bin/bash
# ncu evidence for the CURRENT build (each capture after the same command exited 0 without ncu): one-step launch list with DRAM
# bytes, and one --set full capture of the one-pass kernel at layer 1.  Results land in gpurun_out/; scripts/ncu_traffic.py and
# scripts/ncu_summary.py turn them into profiles/r2_ncu_traffic.json / profiles/r2_ncu_onepassL1.txt.
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import bench; print(bench.lib_sha16())" > $OUT/r2_capture_lib_sha16.txt      # source hash of the library (bench.lib_sha16)
timeout 300 python benchmarks/profile_step.py > $OUT/profstep_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file $OUT/r2_step_launches.csv python benchmarks/profile_step.py > $OUT/profstep_ncu.log 2>&1
echo "launch list exit $?"
timeout 300 python benchmarks/onepass_ab.py --layers 1 --iters 2 > $OUT/onepass_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ldconv_onepass -s 1 -c 1 -f -o $OUT/prof_onepassL1 \
    python benchmarks/onepass_ab.py --layers 1 --iters 2 > $OUT/onepass_ncu.log 2>&1
echo "set-full exit $?"
