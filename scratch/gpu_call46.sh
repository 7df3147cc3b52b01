#!/bin/bash
mkdir -p gpurun_out
python scratch/fn_host_time.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_flownet2.py tests/test_flow_ops_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -2
