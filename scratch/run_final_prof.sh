#!/bin/bash
# round-end evidence: per-launch metrics of one train step + full captures of the three tensor-tile kernels
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r01_v5_step_metrics.csv python scratch/one_step.py > gpurun_out/ncu_step.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_igemm_flat|k_wgrad_flat" --launch-skip 1 -c 3 -o gpurun_out/r01_v5_flat -f python scratch/one_step.py > gpurun_out/ncu_full1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_igemm_tc3" --launch-skip 5 -c 2 -o gpurun_out/r01_v5_tc3 -f python scratch/one_step.py > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
