#!/bin/bash
N=8
mkdir -p gpurun_out
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu $1 2>gpurun_out/scale8.err | tail -1 > gpurun_out/scale8_$2.json
python -c "
import json
d=json.loads(open('gpurun_out/scale8_$2.json').read()); print('N=$N $1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['details']['grad_exchange'])"
}
run "" sharded
run "--no-shard-optimizer" allreduce
run "" sharded2
