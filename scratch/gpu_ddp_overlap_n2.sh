#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q --timeout 800 2>&1 | tail -8 > gpurun_out/tests13.log
tail -3 gpurun_out/tests13.log
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu $2 2>gpurun_out/bench13.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
}
for c in 2 4 8 16; do
NCCL_MAX_CTAS=$c NCCL_MIN_CTAS=1 run "ctas$c" ""
NCCL_MAX_CTAS=$c NCCL_MIN_CTAS=1 run "ctas$c" "--no-overlap"
done
run "default" ""
run "default" "--no-overlap"
