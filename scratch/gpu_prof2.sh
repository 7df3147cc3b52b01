#!/bin/bash
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
VV_PREC=2 timeout 600 ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/r02_f16_step_warm2.csv python scratch/one_step2.py > gpurun_out/ncu_step_warm2.log 2>&1
VV_PREC=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bn_bwd_apply" --launch-skip 30 -c 2 -o gpurun_out/r02_bn_bwd_apply -f python scratch/one_step2.py > gpurun_out/ncu_full_bn.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
