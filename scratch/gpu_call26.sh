#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do VECVAD_TILE_FIRST=$v python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench26.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tile_first $v', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']))"; done
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q --timeout 500 -x 2>&1 | tail -2
