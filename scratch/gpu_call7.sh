#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 10 --precision f16 --no-cpu > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
VECVAD_PREP_SIDE=1 python bench.py --steps 50 --warmup 10 --precision f16 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python bench.py --steps 50 --warmup 10 --precision f16 --no-cpu > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
for f in a b c; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']))
except Exception as e:
    print('$f', 'ERR', e, open('gpurun_out/bench_$f.err').read()[-800:])
PY
done
