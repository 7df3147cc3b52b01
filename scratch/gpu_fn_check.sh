#!/bin/bash
# FlowNet2 side only: parity tests, then the stack's time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q -x --timeout 600 2>&1 | tail -2
timeout 300 python bench_flow.py --flownet2 --iters 20 2>&1 | tail -1 | cut -c1-140
