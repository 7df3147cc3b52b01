#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_auroc_parity_gpu.py -m gpu -q --timeout 1400 -x 2>&1 | tail -30 > gpurun_out/tests17.log
tail -30 gpurun_out/tests17.log
