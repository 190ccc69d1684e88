#!/bin/bash
# final evidence for the CURRENT library sources: launch list + one-pass set-full (r2_capture.sh), a set-full capture of EVERY launch of
# one step, the layer-1 backward launch list, bench line
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $OUT/r2_final_pytest.log
bash scripts/r2_capture.sh
timeout 900 ncu --profile-from-start off --set full --clock-control none -f -o $OUT/prof_step_full python benchmarks/profile_step.py > $OUT/profstep_full.log 2>&1
echo "step set-full exit $?"; ls -la $OUT/prof_step_full.ncu-rep | awk '{print $5}'
timeout 300 python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/r2_bwd_launches_L1.csv python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_ncu.log 2>&1
echo "bwd launch list exit $?"
python benchmarks/profile_step_insitu.py --per-launch 2>/dev/null > $OUT/r2_final_step_insitu.txt; head -3 $OUT/r2_final_step_insitu.txt
python benchmarks/profile_train.py --batch 128 > $OUT/r2_profile_train_b128.txt 2>/dev/null; head -3 $OUT/r2_profile_train_b128.txt | cut -c1-120
