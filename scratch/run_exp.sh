#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_fuse1.json 2> gpurun_out/bench_fuse1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_fuse1.json')); print(d['ms_per_step'], d['value'], d['config']['final_losses'], d['kernel_classes_ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
