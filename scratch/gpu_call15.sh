#!/bin/bash
mkdir -p gpurun_out
python scratch/fn_layer_times.py > gpurun_out/fn_layers.txt 2>&1
head -50 gpurun_out/fn_layers.txt
timeout 600 python -m pytest tests/test_flownet2.py -m gpu -q --timeout 600 -k calc_optical 2>&1 | tail -5
