#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_unet_gpu.py -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/tests6.log
python bench.py --steps 50 --warmup 10 --precision f16 --no-cpu > gpurun_out/bench_f16_v4.json 2> gpurun_out/bench_f16_v4.err
tail -4 gpurun_out/tests6.log
for f in f16_v4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['config']['final_losses'])
except Exception as e:
    print('$f', 'ERR', e, open('gpurun_out/bench_$f.err').read()[-800:])
PY
done
