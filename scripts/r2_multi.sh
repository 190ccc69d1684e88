#!/bin/bash
# multi-GPU legs of bench.py the way the driver launches them: weak scaling (64 images per GPU) incl. the config-4 training leg, and
# strong scaling (one 64-image batch split over the ranks).  usage: bash scripts/r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 \
    > gpurun_out/r2_bench_${N}gpu_weak.json 2> gpurun_out/r2_bench_${N}gpu_weak.err; tail -2 gpurun_out/r2_bench_${N}gpu_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong \
    > gpurun_out/r2_bench_${N}gpu_strong.json 2> gpurun_out/r2_bench_${N}gpu_strong.err; tail -2 gpurun_out/r2_bench_${N}gpu_strong.err
python - <<PY
import json
for mode in ("weak", "strong"):
    try:
        d=json.load(open('gpurun_out/r2_bench_${N}gpu_%s.json' % mode))
    except Exception as e:
        print(mode, 'no line', e); continue
    print(mode, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, d['config']['per_gpu_batch'], 'e2e', d['e2e']['value'], d['e2e'].get('raw_head_output',{}).get('value'),
          'train', {k:d.get('config4_train',{}).get(k) for k in ('value','ms_per_step','error')})
PY
