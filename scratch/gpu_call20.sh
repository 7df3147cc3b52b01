#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_unet_gpu.py tests/test_parity_b128_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -8 > gpurun_out/tests20.log
tail -4 gpurun_out/tests20.log
for i in 1 2; do python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary > gpurun_out/bench20_$i.json 2> gpurun_out/bench20_$i.err; done
python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary --precision tf32 > gpurun_out/bench20_tf32.json 2>/dev/null
python bench.py --steps 30 --warmup 10 --no-cpu --no-secondary --net full > gpurun_out/bench20_full.json 2>/dev/null
for f in 1 2 tf32 full; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench20_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['details']['final_losses'])
except Exception as e:
    print('$f', 'ERR', e)
PY
done
VECVAD_FLAT_TRACE=1 VECVAD_TC3_TRACE=1 VV_STEPS=2 python scratch/one_step.py 2>&1 | grep trace | tail -37 > gpurun_out/trace20.txt
head -3 gpurun_out/trace20.txt | cut -c1-300
