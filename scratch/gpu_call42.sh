#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fn_conv -c 4 -o gpurun_out/fn_conv_r02b -f python scratch/fn_one_layer.py once > gpurun_out/ncu42.log 2>&1
tail -2 gpurun_out/ncu42.log
