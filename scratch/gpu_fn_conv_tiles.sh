#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q -x --timeout 600 2>&1 | tail -3
echo auto; python scratch/fn_one_layer.py 2>&1 | tail -8
for t in 0 1 2 3 4 5 6; do echo TILE $t; VECVAD_FN_TILE=$t python scratch/fn_one_layer.py 2>&1 | tail -8; done
timeout 300 python bench_flow.py --flownet2 --iters 10 2>&1 | tail -1 | cut -c1-200
