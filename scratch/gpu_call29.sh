#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_bn_apply|k_bn_bwd" --launch-skip 42 -c 44 -o gpurun_out/r02_bn_final -f python scratch/one_step.py > gpurun_out/ncu_bn.log 2>&1
ls -la gpurun_out/r02_bn_final.ncu-rep
