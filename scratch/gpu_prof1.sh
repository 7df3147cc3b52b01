#!/bin/bash
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
VV_PREC=2 timeout 600 ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/r02_f16_step_warm.csv python scratch/one_step2.py > gpurun_out/ncu_step_warm.log 2>&1
tail -3 gpurun_out/ncu_step_warm.log; wc -l gpurun_out/r02_f16_step_warm.csv
