#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_conv_gpu.py tests/test_parity_b128_gpu.py -m gpu -q --timeout 800 -x 2>&1 | tail -3
for v in 1 2; do python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench32.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('run $v', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['details']['final_losses'])"; done
