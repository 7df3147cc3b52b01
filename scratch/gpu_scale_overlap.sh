#!/bin/bash
# N GPUs: phased (overlapped) gradient exchange against one all-reduce after the backward, by NCCL CTA limit; $1 = N
N=$1
mkdir -p gpurun_out
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu $2 2>gpurun_out/scale_ov.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N $1 $2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
}
run "default" ""
run "default" "--overlap"
NCCL_MAX_CTAS=16 NCCL_MIN_CTAS=1 run "ctas16" "--overlap"
NCCL_MAX_CTAS=8 NCCL_MIN_CTAS=1 run "ctas8" "--overlap"
NCCL_MAX_CTAS=32 NCCL_MIN_CTAS=1 run "ctas32" "--overlap"
