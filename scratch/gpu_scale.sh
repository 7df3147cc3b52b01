#!/bin/bash
# weak-scaling bench lines at N GPUs (the driver's launch line); $1 = N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 10 2>gpurun_out/scale_n$N.err | tail -1 > gpurun_out/r02_bench_final_n$N.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final_n$N.json').read()); print('N=$N', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['details']['grad_exchange'], d['clocks'])"
