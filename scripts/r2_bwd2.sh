#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "offset_conv_backward or accumulator or train" 2>&1 | tail -3
timeout 300 python benchmarks/one_bwd.py --layer 1 > gpurun_out/onebwd_plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r2_bwd_launches_L1.csv python benchmarks/one_bwd.py --layer 1 > gpurun_out/onebwd_ncu.log 2>&1
echo "launch list exit $?"
python bench.py --steps 5 --warmup 3 --train-steps 6 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d['config4_train'].get(k) for k in ('value','ms_per_step','final_loss','error')})"
