#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q --timeout 600 2>&1 | tail -12 > gpurun_out/tests16.log
tail -6 gpurun_out/tests16.log
python scratch/fn_layer_times.py > gpurun_out/fn_layers2.txt 2>&1
head -30 gpurun_out/fn_layers2.txt
timeout 300 python bench_flow.py --flownet2 --iters 10 2>&1 | tail -2
