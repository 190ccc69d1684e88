#!/bin/bash
python benchmarks/profile_train_copies.py 2>/dev/null | head -45
