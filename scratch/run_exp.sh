#!/bin/bash
for s in 2 3 4; do
  VECVAD_FLAT_STAGES=$s VV_TIME=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_igemm_flat --log-file gpurun_out/flat16_times_s$s.csv python scratch/flat16_probe.py > /dev/null 2>&1
done
VECVAD_FLAT_TRACE=1 VECVAD_FLAT_STAGES=2 VV_TIME=1 timeout 100 python scratch/flat16_probe.py 2>&1 | grep "flat trace" | awk 'NR%3==0'
VECVAD_FLAT_TRACE=1 VECVAD_FLAT_STAGES=4 VV_TIME=1 timeout 100 python scratch/flat16_probe.py 2>&1 | grep "flat trace" | awk 'NR%3==0'
