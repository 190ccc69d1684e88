#!/bin/bash
# closing evidence of session 4: full GPU tests, smoke, both bench arms, one-step launch list, in-situ step profile
set -u
OUT=gpurun_out; TAG=s4b
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cut -c1-200 $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref exit $?"; cut -c1-160 $OUT/bench_ref_$TAG.json
timeout 300 python benchmarks/profile_step_insitu.py > $OUT/step_insitu_$TAG.txt 2>&1; head -14 $OUT/step_insitu_$TAG.txt | tail -12
timeout 300 python benchmarks/profile_step.py > $OUT/profstep_plain_$TAG.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $OUT/step_launches_$TAG.csv python benchmarks/profile_step.py > $OUT/profstep_ncu_$TAG.log 2>&1
echo "ncu exit $?"
