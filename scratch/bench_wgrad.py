"""Microbenchmark of the weight-gradient tiles (use_tc 2 = tap-reuse tcgen05 tiles)."""
import os, sys, torch
sys.path.insert(0, '.')
from vec_vad_b200 import _lib
def runw(b, h, cin, cout, use_tc, iters=int(os.environ.get('VV_ITERS', '30'))):
    x = torch.randn(b, h, h, cin, device='cuda'); go = torch.randn(b, h, h, cout, device='cuda')
    dw = torch.empty(cout, cin, 3, 3, device='cuda'); scratch = torch.empty(9*cout*cin, device='cuda')
    L = _lib.lib()
    def f():
        _lib.check(L.vecvad_conv3x3_wgrad(_lib.ptr(x), cin, _lib.ptr(go), _lib.ptr(dw), _lib.ptr(scratch), b, h, h, cin, cout, use_tc, _lib.cur_stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * b * h * h * cout * cin * 9
    print('WGRAD B=%d H=%d %d->%d tc=%d: %.1f us  %.1f TFLOP/s' % (b, h, cin, cout, use_tc, t, fl / t / 1e6), flush=True)
NB = int(os.environ.get('VV_B', '768'))
MODES = [int(m) for m in os.environ.get('VV_MODES', '2').split(',')]
for cfg in [(NB, 32, 32, 32), (NB, 32, 64, 32), (NB, 16, 64, 64), (NB, 16, 128, 64), (NB, 8, 128, 128), (NB, 8, 256, 128), (NB, 4, 256, 256)]:
    for tc in MODES:
        runw(*cfg, tc)
