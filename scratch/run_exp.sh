#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 0 1; do VECVAD_WGRAD_FLAT=$m timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_wgf$m.json 2> gpurun_out/bench_wgf$m.err; done
python -c "
import json
for m in (0,1):
    d=json.load(open('gpurun_out/bench_wgf%d.json'%m)); print('WGRAD_FLAT=%d'%m, d['ms_per_step'], d['value'], d['config']['final_losses'], d['kernel_classes_ms_per_step'], d['e2e']['ms_per_step'])"
