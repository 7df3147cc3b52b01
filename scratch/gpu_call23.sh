#!/bin/bash
mkdir -p gpurun_out
for p in 1 2 3; do VECVAD_NET_PARTS=$p python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench23.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('parts $p', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']))"; done
VECVAD_WGRAD_STREAM=0 python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench23.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no side stream', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'])"
