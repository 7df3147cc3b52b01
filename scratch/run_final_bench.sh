#!/bin/bash
python bench.py > gpurun_out/r01_bench_v5_n1.json 2> gpurun_out/r01_bench_v5_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_v5_reference.json 2>/dev/null
python bench.py --net full --batch 128 --steps 60 --warmup 5 --no-cpu > gpurun_out/r01_bench_v5_full.json 2>/dev/null
python bench_flow.py > gpurun_out/r01_bench_flow_v3.jsonl 2>/dev/null
python bench_flow.py --batch 8 >> gpurun_out/r01_bench_flow_v3.jsonl 2>/dev/null
cut -c1-250 gpurun_out/r01_bench_v5_n1.json; echo; cut -c1-200 gpurun_out/r01_bench_v5_reference.json; echo; cut -c1-200 gpurun_out/r01_bench_v5_full.json; echo; cut -c1-220 gpurun_out/r01_bench_flow_v3.jsonl
