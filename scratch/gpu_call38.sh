#!/bin/bash
mkdir -p gpurun_out
python scratch/fn_one_layer.py 2>&1 | tail -6
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fn_conv -c 8 -o gpurun_out/fn_conv_r02 -f python scratch/fn_one_layer.py once > gpurun_out/ncu38.log 2>&1
tail -3 gpurun_out/ncu38.log; ls -la gpurun_out/*.ncu-rep
