#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 500 -x 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary > gpurun_out/bench21_$i.json 2> gpurun_out/bench21_$i.err; done
for f in 1 2; do python - <<PY
import json
d=json.loads(open('gpurun_out/bench21_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']))
PY
done
VECVAD_FLAT_TRACE=1 VV_STEPS=2 python scratch/one_step.py 2>&1 | grep "flat trace" | tail -8 | cut -c1-330
