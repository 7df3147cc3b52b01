import sys, torch
sys.path.insert(0, '.')
exec(open('scratch/bench_conv.py').read().split("for cfg in")[0])
run(768, 32, 32, 32, 2)
run(768, 8, 256, 128, 2)
