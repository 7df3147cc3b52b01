#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_crop_resize.py tests/test_pipeline.py tests/test_unet_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/tests11.log
tail -6 gpurun_out/tests11.log
