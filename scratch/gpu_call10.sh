#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 10 > gpurun_out/bench10.json 2> gpurun_out/bench10.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
    for s in d.get('secondary', []): print('  ', s.get('workload','')[:60], round(s.get('value',0)), s.get('roofline',{}).get('frac'), s.get('us_per_launch'), s.get('error'))
except Exception as e:
    print('ERR', e, open('gpurun_out/bench10.err').read()[-1500:])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_flow_launches_b8.csv -k regex:"k_corr|k_resample|k_parity|k_channelnorm" python bench_flow.py --batch 8 --iters 2 --no-reference > gpurun_out/ncu_flow.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_flow_launches_b1.csv -k regex:"k_corr|k_resample|k_parity|k_channelnorm" python bench_flow.py --batch 1 --iters 2 --no-reference >> gpurun_out/ncu_flow.log 2>&1
grep -v '^==' gpurun_out/r02_flow_launches_b8.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if int(r['ID'])>=18: print(r['ID'], r['Kernel Name'][:40], r['Grid Size'], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
" | tail -40
