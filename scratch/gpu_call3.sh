#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 10 --precision tf32 --no-cpu > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
python bench.py --steps 50 --warmup 10 --precision f16 --no-cpu > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
VECVAD_WGRAD_STREAM=0 python bench.py --steps 50 --warmup 10 --precision f16 --no-cpu > gpurun_out/bench_f16_1s.json 2> gpurun_out/bench_f16_1s.err
python bench.py --steps 30 --warmup 10 --precision f16 --no-cpu --net full > gpurun_out/bench_f16_full.json 2> gpurun_out/bench_f16_full.err
for f in tf32 f16 f16_1s f16_full; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['config']['final_losses'])
except Exception as e:
    print('$f', 'ERR', e, open('gpurun_out/bench_$f.err').read()[-800:])
PY
done
