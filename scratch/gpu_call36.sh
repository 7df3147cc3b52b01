#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flow_ops_gpu.py tests/test_unet_gpu.py tests/test_parity_b128_gpu.py -m gpu -q -x --timeout 800 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 100 --warmup 20 --no-cpu --no-secondary 2>gpurun_out/b36.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['roofline']['frac'])"
done
python bench_flow.py 2>&1 | tail -4 | cut -c1-400; python bench_flow.py --batch 8 2>&1 | tail -4 | cut -c1-400
