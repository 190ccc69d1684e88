#!/bin/bash
python benchmarks/profile_train.py 2>/dev/null | head -45
