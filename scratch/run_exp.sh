#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"k_bn|k_maxpool|k_colsum" --log-file gpurun_out/bn_list.csv python scratch/one_step.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[l for l in open('gpurun_out/bn_list.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(rows):
    n=r['Kernel Name'].split('(')[0]; v=float(r['Metric Value'].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}.get(r['Metric Unit'],1e-3)
    agg[n][0]+=1; agg[n][1]+=v
for k,v in agg.items(): print(k, v)
PY
timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value']), d['kernel_classes_ms_per_step'], d['e2e']['ms_per_step'])"
