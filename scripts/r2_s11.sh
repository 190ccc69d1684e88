#!/bin/bash
for l in 15 1; do python benchmarks/trace_onepass.py --layer $l 2>&1 | tail -60; done
