#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_b128.jsonl
timeout 1200 python -m pytest tests/test_parity_b128_gpu.py -q --timeout 900 2>&1 | tail -40 > gpurun_out/parity8.log
python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/bench8.json 2> gpurun_out/bench8.err
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_parity_b128_gpu.py 2>&1 | tail -15 > gpurun_out/tests8.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scratch/sanity_small.py > gpurun_out/racecheck8.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/racecheck8.txt
tail -5 gpurun_out/parity8.log; tail -3 gpurun_out/tests8.log; tail -5 gpurun_out/racecheck8.txt
cat gpurun_out/parity_b128.jsonl
python - <<PY
import json
d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'])
PY
