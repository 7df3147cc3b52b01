#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/tests14.log
tail -25 gpurun_out/tests14.log
timeout 300 python bench_flow.py --flownet2 --iters 5 2>&1 | tail -3
