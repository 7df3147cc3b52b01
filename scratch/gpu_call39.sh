#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q -x --timeout 600 2>&1 | tail -3
echo auto; python scratch/fn_one_layer.py 2>&1 | tail -5
echo TN64; VECVAD_FN_TN=64 python scratch/fn_one_layer.py 2>&1 | tail -5
echo TN128; VECVAD_FN_TN=128 python scratch/fn_one_layer.py 2>&1 | tail -5
python scratch/fn_layer_times.py > gpurun_out/fn_layers4.txt 2>&1
head -3 gpurun_out/fn_layers4.txt
timeout 300 python bench_flow.py --flownet2 --iters 10 2>&1 | tail -1 | cut -c1-200
