#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke(), reference arm, our arm
mkdir -p gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/rehearsal_tests.log 2>&1
tail -4 gpurun_out/rehearsal_tests.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/rehearsal_smoke.log 2>&1
tail -12 gpurun_out/rehearsal_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/rehearsal_ref.log 2>&1
tail -4 gpurun_out/rehearsal_ref.log | cut -c1-300
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/rehearsal_bench.log 2>&1
tail -5 gpurun_out/rehearsal_bench.log | cut -c1-600
