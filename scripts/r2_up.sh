#!/bin/bash
python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --train-steps 6 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d['config4_train'].get(k) for k in ('value','ms_per_step','final_loss','error')})"
