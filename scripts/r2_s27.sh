#!/bin/bash
python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --train-steps 0 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'], d['e2e']['value'], d['e2e']['raw_head_output']['value'])"
