#!/bin/bash
for c in 1.0 0.6 0.35 0.15; do echo split cost x$c; VECVAD_FN_SPLIT_COST=$c timeout 120 python bench_flow.py --flownet2 --iters 20 2>&1 | tail -1 | cut -c1-100; done
