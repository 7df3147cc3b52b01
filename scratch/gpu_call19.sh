#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flow_ops_gpu.py -m gpu -q --timeout 800 2>&1 | tail -5 > gpurun_out/tests19.log
tail -3 gpurun_out/tests19.log
for b in 8 1; do for v in "X=0" "VECVAD_RESAMPLE_GENERIC=1" "VECVAD_RESAMPLE_FP32=1"; do echo "# $v batch $b"; env $v python bench_flow.py --batch $b --iters 20 --no-reference | grep -v correlation; done; done > gpurun_out/bench_flow19.jsonl 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_flow19.jsonl'):
    try:
        d=json.loads(l); print(d['op'], d['batch'], round(d['us'],1), round(d['frac_of_hbm_peak'],3))
    except Exception: print(l.strip()[:200])
PY
