"""Microbenchmark: persistent per-dx-box tiles (use_tc 2) against their pair variant (use_tc 4) below 32x32 resolution."""
import os, sys, torch
sys.path.insert(0, '.')
from vec_vad_b200 import _lib
def run(b, h, cin, cout, use_tc, iters=int(os.environ.get('VV_ITERS', '30'))):
    x = torch.randn(b, h, h, cin, device='cuda'); w = torch.randn(cout, cin, 3, 3, device='cuda'); bias = torch.randn(cout, device='cuda')
    out = torch.empty(b, h, h, cout, device='cuda'); stats = torch.zeros(2*cout, dtype=torch.float64, device='cuda'); scratch = torch.empty(9*cout*cin, device='cuda')
    L = _lib.lib()
    def f():
        _lib.check(L.vecvad_conv3x3_forward(_lib.ptr(x), cin, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(stats), _lib.ptr(scratch), b, h, h, cin, cout, use_tc, _lib.cur_stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * b * h * h * cout * cin * 9
    print('B=%d H=%d %d->%d tc=%d: %.1f us  %.1f TFLOP/s' % (b, h, cin, cout, use_tc, t, fl / t / 1e6), flush=True)
NB = int(os.environ.get('VV_B', '768'))
for cfg in [(NB, 16, 32, 64), (NB, 16, 64, 64), (NB, 16, 128, 64), (NB, 8, 64, 128), (NB, 8, 128, 128), (NB, 8, 256, 128), (NB, 4, 128, 256), (NB, 4, 256, 256)]:
    for tc in (2, 4):
        run(*cfg, tc)
