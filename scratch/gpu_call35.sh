#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q --timeout 800 2>&1 | tail -3
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu $1 2>gpurun_out/scale_sh.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N $1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['details']['final_losses'])"
}
run ""
run "--shard-optimizer"
run ""
run "--shard-optimizer"
