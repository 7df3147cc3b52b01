#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_pipeline.py tests/test_auroc_parity_gpu.py -m gpu -q --timeout 800 -x 2>&1 | tail -3
for v in 0 1 0 1; do VECVAD_TAIL_ADAM=$v python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench33.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tail_adam $v', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['details']['final_losses'])"; done
