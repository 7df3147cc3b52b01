#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q --timeout 800 2>&1 | tail -15 > gpurun_out/tests12.log
tail -6 gpurun_out/tests12.log
for mode in "" "--no-overlap" "" "--no-overlap"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu $mode 2>gpurun_out/bench12.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mode', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['details']['grad_exchange'])"
done
python bench.py --steps 50 --warmup 10 --no-cpu --no-secondary | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
