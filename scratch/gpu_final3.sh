#!/bin/bash
# last refresh of the FlowNet2-side evidence and the bench line on the final build
mkdir -p gpurun_out
for b in 1 8; do python bench_flow.py --batch $b --iters 20; done > gpurun_out/r02_bench_flow_final.jsonl 2>&1
python bench_flow.py --flownet2 --iters 20 >> gpurun_out/r02_bench_flow_final.jsonl 2>&1
python scratch/fn_layer_times.py > gpurun_out/r02_flownet2_layers.txt 2>&1
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_final_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
for s in d.get('secondary', []): print('  ', s.get('workload','')[:70], round(s.get('value',0),1), s.get('roofline',{}).get('frac'), s.get('error'))
PY
grep -h 'flownet2_forward' gpurun_out/r02_bench_flow_final.jsonl | cut -c1-150
head -1 gpurun_out/r02_flownet2_layers.txt
