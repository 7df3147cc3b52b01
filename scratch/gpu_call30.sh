#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q --timeout 600 2>&1 | tail -3
python scratch/fn_layer_times.py > gpurun_out/fn_layers3.txt 2>&1
head -12 gpurun_out/fn_layers3.txt
timeout 300 python bench_flow.py --flownet2 --iters 10 2>&1 | tail -1 | cut -c1-200
