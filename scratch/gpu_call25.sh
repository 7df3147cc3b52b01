#!/bin/bash
mkdir -p gpurun_out
for v in 148 222 296 444 592; do VECVAD_WG_TARGET=$v python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench25.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('target $v', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']))"; done
