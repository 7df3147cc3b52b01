#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_conv_gpu.py tests/test_pipeline.py -m gpu -q --timeout 800 -x 2>&1 | tail -3
for v in 1 2; do python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary 2>gpurun_out/bench27.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('run $v', round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['details']['final_losses'])"; done
M=gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/glue27.csv -k regex:"k_outconv|k_prep|k_adam|k_cubes|k_losses|k_scatter|k_maxpool|k_colsum" python scratch/one_step.py > /dev/null 2>&1
grep -v "^==" gpurun_out/glue27.csv | python -c "
import csv,sys,collections
a=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(sys.stdin):
    n=r['Kernel Name'].split('(')[0].replace('void ','').replace('<unnamed>::','')
    a[n][0]+=1; a[n][1]+=float(r['Metric Value'].replace(',',''))/1e3
for k,v in sorted(a.items(), key=lambda kv:-kv[1][1]): print('%-36s %3d launches (2 steps) %8.1f us per step'%(k[:36], v[0], v[1]/2))
"
