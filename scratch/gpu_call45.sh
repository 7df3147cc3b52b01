#!/bin/bash
mkdir -p gpurun_out
for mb in 16 64 256; do echo scratch $mb MB; VECVAD_FN_SCRATCH_MB=$mb timeout 300 python bench_flow.py --flownet2 --iters 10 2>&1 | tail -1 | cut -c1-120; done
timeout 900 python -m pytest tests/test_flownet2.py -m gpu -q -x --timeout 600 2>&1 | tail -2
