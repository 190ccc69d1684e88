#!/bin/bash
python benchmarks/trace_zc.py --cin 32 --cout 96 --hw 160 2>/dev/null | head -44
