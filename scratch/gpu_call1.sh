#!/bin/bash
# round 2, GPU call 1: reference-kernel fixtures, full gpu suite, tf32 peak, baseline bench + flow bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_b128.jsonl
python tests/golden/make_flow_golden.py gpurun_out/flow_ops.npz > gpurun_out/make_flow_golden.log 2>&1
cp gpurun_out/flow_ops.npz tests/golden/flow_ops.npz
python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -m pytest tests/test_flow_reference_golden.py -q > gpurun_out/pytest_flow_cpu.log 2>&1
python scratch/measure_tf32_peak.py gpurun_out/tf32_peak.json > gpurun_out/tf32_peak.log 2>&1
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r02_base.json 2> gpurun_out/bench_r02_base.err
for b in 1 8; do python bench_flow.py --batch $b; done > gpurun_out/bench_flow_r02_base.jsonl 2> gpurun_out/bench_flow_r02_base.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/tf32_peak.log | tail -2; cat gpurun_out/make_flow_golden.log | tail -3; tail -3 gpurun_out/pytest_flow_cpu.log
