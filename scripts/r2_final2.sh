#!/bin/bash
# final single-GPU evidence for the current build (round 2, second half): whole GPU suite, bench line, LDConv A/B, in-situ step profile,
# ncu launch lists + set-full captures (one-pass kernel, first-layer kernel, backward scatter), backward / training profiles
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $OUT/r2_final_pytest.log
python bench.py --steps 30 --warmup 5 > $OUT/r2_final_bench.json 2> $OUT/r2_final_bench.err; tail -2 $OUT/r2_final_bench.err
python benchmarks/onepass_ab.py --iters 7 > $OUT/r2_final_onepass_ab.jsonl 2>/dev/null; tail -1 $OUT/r2_final_onepass_ab.jsonl
python benchmarks/profile_step_insitu.py --per-launch 2>/dev/null > $OUT/r2_final_step_insitu.txt; head -9 $OUT/r2_final_step_insitu.txt
bash scripts/r2_capture.sh
# first layer: A/B + one set-full capture
python benchmarks/l0_ab.py --iters 7 > $OUT/r2_l0_ab.jsonl 2>/dev/null; cat $OUT/r2_l0_ab.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l0_tc -s 2 -c 1 -f -o $OUT/prof_l0tc python benchmarks/l0_ab.py --iters 1 > $OUT/l0tc_ncu.log 2>&1
echo "l0 set-full exit $?"
# backward of one LDConv layer (layer 1, batch 64): launch list + set-full capture of the scatter
timeout 300 python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/r2_bwd_launches_L1.csv python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_ncu.log 2>&1
echo "bwd launch list exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scatter_bwd_tiled2 -c 1 -f -o $OUT/prof_scatterL1_tiled \
    python benchmarks/one_bwd.py --layer 1 > $OUT/onebwd_ncu2.log 2>&1
echo "scatter set-full exit $?"
python benchmarks/acc16_error.py > $OUT/r2_acc16_error.jsonl 2>/dev/null; tail -2 $OUT/r2_acc16_error.jsonl
python benchmarks/ldconv_layers.py --bwd --iters 5 2>/dev/null | grep -E "gather_bwd|gemm_bwd" > $OUT/r2_scatter_layers.jsonl; grep bf16_acc $OUT/r2_scatter_layers.jsonl | head -3 | cut -c1-170
python benchmarks/profile_train.py --batch 128 > $OUT/r2_profile_train_b128.txt 2>/dev/null; head -4 $OUT/r2_profile_train_b128.txt | cut -c1-120
python benchmarks/profile_train_copies.py --batch 128 > $OUT/r2_train_copies_b128.txt 2>/dev/null; head -2 $OUT/r2_train_copies_b128.txt | cut -c1-120
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_final_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches_per_step','lib_sha16')})
print('e2e', d['e2e']['value'], d['e2e'].get('raw_head_output'))
print('train', {k:d.get('config4_train',{}).get(k) for k in ('value','ms_per_step','error')})
r=d['roofline']; print('roof', r['achieved'], r['frac'], r['us_per_launch'], r['all_onepass_launches'])
print('scatter', {k:d['roofline_scatter'].get(k) for k in ('achieved','frac','us_per_step','largest_launch')})
PY
