#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q --timeout 300 2>&1 | tail -40 > gpurun_out/conv_tests.log
timeout 900 python -m pytest tests/test_unet_gpu.py -q --timeout 600 2>&1 | tail -40 > gpurun_out/unet_tests.log
timeout 600 python scratch/grad_diag.py net4_flow_b2 0 1 2 2>&1 | grep -v "^   " > gpurun_out/grad_diag.log
tail -25 gpurun_out/conv_tests.log; tail -25 gpurun_out/unet_tests.log; cat gpurun_out/grad_diag.log
