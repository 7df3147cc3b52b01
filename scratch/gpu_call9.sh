#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_b128.jsonl
VECVAD_PDL=0 python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/bench9_pdl0.json 2> gpurun_out/bench9_pdl0.err
VECVAD_PDL=1 python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/bench9_pdl1.json 2> gpurun_out/bench9_pdl1.err
VECVAD_PDL=0 python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/bench9_pdl0b.json 2> gpurun_out/bench9_pdl0b.err
VECVAD_PDL=1 python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/bench9_pdl1b.json 2> gpurun_out/bench9_pdl1b.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 > gpurun_out/tests9.log
tail -4 gpurun_out/tests9.log
for f in pdl0 pdl1 pdl0b pdl1b; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench9_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['config']['final_losses'])
except Exception as e:
    print('$f', 'ERR', e, open('gpurun_out/bench9_$f.err').read()[-800:])
PY
done
