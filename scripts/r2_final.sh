#!/bin/bash
# final evidence for the CURRENT library sources: launch list + one-pass set-full (r2_capture.sh), the layer-1 backward launch list, bench line
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $OUT/r2_final_pytest.log
bash scripts/r2_capture.sh
# stamp the capture on the box and time the bench line against it (one call instead of two)
python scripts/ncu_traffic.py > /dev/null && cp profiles/r2_ncu_traffic.json $OUT/r2_ncu_traffic.json
python bench.py --steps 30 --warmup 5 > $OUT/r2_final_bench.json 2> $OUT/r2_final_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_final_bench.json'))
print(d['value'], d['e2e']['value'], d['lib_sha16'], d['roofline']['traffic'], d.get('step_roofline',{}).get('frac'), d['config4_train']['ms_per_step'], d['roofline_scatter']['frac'])"
timeout 300 python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/r2_bwd_launches_L1.csv python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_ncu.log 2>&1
echo "bwd launch list exit $?"
python benchmarks/profile_step_insitu.py --per-launch 2>/dev/null > $OUT/r2_final_step_insitu.txt; head -3 $OUT/r2_final_step_insitu.txt
python benchmarks/profile_train.py --batch 128 > $OUT/r2_profile_train_b128.txt 2>/dev/null; head -3 $OUT/r2_profile_train_b128.txt | cut -c1-120
