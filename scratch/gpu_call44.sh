#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q -x --timeout 600 2>&1 | tail -2
python scratch/fn_one_layer.py 2>&1 | tail -8
python scratch/fn_layer_times.py > gpurun_out/fn_layers6.txt 2>&1
head -1 gpurun_out/fn_layers6.txt
timeout 300 python bench_flow.py --flownet2 --iters 10 2>&1 | tail -1 | cut -c1-200
