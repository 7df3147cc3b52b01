#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flow_ops_gpu.py tests/test_flownet2.py tests/test_auroc_parity_gpu.py -m gpu -q --timeout 800 2>&1 | tail -12 > gpurun_out/tests18.log
tail -6 gpurun_out/tests18.log
for b in 1 8; do python bench_flow.py --batch $b --iters 20 --no-reference; VECVAD_RESAMPLE_TILED=0 python bench_flow.py --batch $b --iters 20 --no-reference | grep -v correlation; done > gpurun_out/bench_flow18.jsonl 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_flow18.jsonl'):
    try:
        d=json.loads(l); print(d['op'], d['batch'], round(d['us'],1), round(d['frac_of_hbm_peak'],3))
    except Exception: print(l.strip()[:200])
PY
cat gpurun_out/auroc_parity.json
