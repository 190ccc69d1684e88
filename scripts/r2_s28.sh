#!/bin/bash
python -m pytest tests/test_gpu_loss.py -q -m gpu 2>&1 | tail -8
python benchmarks/profile_loss.py 2>/dev/null | head -2
python benchmarks/train_step.py --batch 128 --steps 6 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', d['value'], d['ms_per_step'], d['final_loss'])"
