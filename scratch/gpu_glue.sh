#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/glue.csv -k regex:"k_outconv|k_prep|k_adam|k_cubes|k_losses|k_scatter|k_maxpool|k_colsum" python scratch/one_step.py > /dev/null 2>&1
grep -v "^==" gpurun_out/glue.csv | python -c "
import csv,sys,collections
a=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(sys.stdin):
    n=r['Kernel Name'].split('(')[0].replace('void ','').replace('<unnamed>::','')
    a[n][0]+=1; a[n][1]+=float(r['Metric Value'].replace(',',''))/1e3
for k,v in sorted(a.items(), key=lambda kv:-kv[1][1]): print('%-36s %3d launches (2 steps) %8.1f us per step'%(k[:36], v[0], v[1]/2))
"
