#!/bin/bash
mkdir -p gpurun_out
python bench_flow.py --no-reference > gpurun_out/flow37_b1.jsonl 2>&1; python bench_flow.py --no-reference --batch 8 > gpurun_out/flow37_b8.jsonl 2>&1
grep -h '"op": "correlation' gpurun_out/flow37_b*.jsonl | cut -c1-260
