#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 1 0; do VECVAD_NO_TC3=$m timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_tc3_no$m.json 2> gpurun_out/bench_tc3_no$m.err; done
python -c "
import json
for m in (1,0):
    d=json.load(open('gpurun_out/bench_tc3_no%d.json'%m)); print('NO_TC3=%d'%m, d['ms_per_step'], d['value'], d['config']['final_losses'], d['kernel_classes_ms_per_step'], d['e2e']['ms_per_step'])"
