#!/bin/bash
run() { env "$@" timeout 60 python scratch/bench_flat.py 2>&1 | grep -v Warning | tail -16; }
timeout 100 python scratch/flat_probe.py 2>&1 | head -7
run
run VECVAD_FLAT_TRACE=1 VV_ITERS=1 | grep prod | awk 'NR%4==0'
for m in 0 1; do VECVAD_FLAT=$m timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_flatv4_$m.json 2> gpurun_out/bench_flatv4_$m.err; done
python -c "
import json
for m in (0,1):
    d=json.load(open('gpurun_out/bench_flatv4_%d.json'%m)); print(m, d['ms_per_step'], d['value'], d['config']['final_losses'], d['kernel_classes_ms_per_step'], d['e2e']['ms_per_step'])"
