#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fn_conv" --launch-skip 130 -c 6 -o gpurun_out/r02_fnconv -f python bench_flow.py --flownet2 --iters 1 > gpurun_out/ncu_fnconv.log 2>&1
ls -la gpurun_out/r02_fnconv.ncu-rep
