#!/bin/bash
VV_PREC=2 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:bn_bwd_apply --launch-skip 15 -c 1 -o gpurun_out/r02_bn_bwd_apply -f python scratch/one_step2.py > gpurun_out/ncu_full_bn1.log 2>&1
VV_PREC=2 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:bn_bwd_reduce --launch-skip 15 -c 1 -o gpurun_out/r02_bn_bwd_reduce -f python scratch/one_step2.py > gpurun_out/ncu_full_bn2.log 2>&1
VV_PREC=2 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bn_apply --launch-skip 14 -c 1 -o gpurun_out/r02_bn_apply -f python scratch/one_step2.py > gpurun_out/ncu_full_bn3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
