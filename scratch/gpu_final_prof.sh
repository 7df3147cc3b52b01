#!/bin/bash
# round-2 evidence with the final build: per-launch metrics of one fp16-operand train step, in-kernel cycle counters, one --set full
# capture of the dominant tile kernel, the bench lines
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_f16_step_metrics.csv python scratch/one_step.py > gpurun_out/ncu_step.log 2>&1
VECVAD_FLAT_TRACE=1 VECVAD_TC3_TRACE=1 VECVAD_WGF_TRACE=1 VV_STEPS=2 python scratch/one_step.py > gpurun_out/r02_f16_trace_raw.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_igemm_tc3" --launch-skip 40 -c 2 -o gpurun_out/r02_tc3_f16 -f python scratch/one_step.py > gpurun_out/ncu_full_tc3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_igemm_flat" --launch-skip 8 -c 1 -o gpurun_out/r02_flat_f16 -f python scratch/one_step.py > gpurun_out/ncu_full_flat.log 2>&1
python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_final_reference.json 2>/dev/null
python bench.py --precision tf32 --steps 50 --no-cpu --no-secondary > gpurun_out/r02_bench_final_tf32.json 2>/dev/null
for b in 1 8; do python bench_flow.py --batch $b --iters 20; done > gpurun_out/r02_bench_flow_final.jsonl 2>&1
python bench_flow.py --flownet2 --iters 10 >> gpurun_out/r02_bench_flow_final.jsonl 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_final_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), d['kernel_classes_ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
for s in d.get('secondary', []): print('  ', s.get('workload','')[:70], round(s.get('value',0),1), s.get('roofline',{}).get('frac'), s.get('error'))
PY
