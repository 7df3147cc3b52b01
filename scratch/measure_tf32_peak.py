"""cuBLAS TF32 matmul throughput on this box, measured the way MEASURED_PEAKS.json's bf16 figure was (torch.matmul 8192^3,
best of 10 = burst, back to back for 4 s = sustained).  The yardstick for kind::tf32 tiles."""
import json
import sys
import time

import torch


def measure(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device='cuda', dtype=dtype)
    b = torch.randn(n, n, device='cuda', dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / k
    fl = 2.0 * n ** 3
    return fl / (best * 1e-3) / 1e12, fl / (sus * 1e-3) / 1e12


if __name__ == '__main__':
    out = {}
    out['tf32_tflops'], out['tf32_tflops_sustained'] = measure(torch.float32, True)
    out['bf16_tflops'], out['bf16_tflops_sustained'] = measure(torch.bfloat16, False)
    out['fp16_tflops'], out['fp16_tflops_sustained'] = measure(torch.float16, False)
    out['how'] = 'torch.matmul 8192^3 (2*N^3 FLOP): best of 10 (burst) and back to back for 4 s (sustained); tf32 = fp32 tensors with allow_tf32'
    out['gpu'] = torch.cuda.get_device_name(0)
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], 'w'), indent=1)
