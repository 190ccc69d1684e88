#!/bin/bash
python benchmarks/profile_loss.py 2>/dev/null | head -40
