"""Times (or, under ncu, just launches once) a few representative FlowNet2 conv layers through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vec_vad_b200 import _lib
SHAPES = [  # cin, cout, k, s, H, W
    (64, 128, 3, 1, 192, 256),
    (64, 128, 5, 2, 192, 256),
    (512, 512, 3, 1, 24, 32),
    (82, 16, 3, 1, 384, 512),
    (12, 64, 7, 2, 384, 512),
    (64, 64, 3, 2, 384, 512),
    (194, 64, 3, 1, 96, 128),
    (162, 32, 3, 1, 192, 256),
]
once = len(sys.argv) > 1 and sys.argv[1] == 'once'
dev = torch.device('cuda:0')
sc = torch.empty(8 << 20, device=dev)
for cin, cout, k, s, H, W in SHAPES:
    x = torch.randn(1, cin, H, W, device=dev); w = torch.randn(cout, cin, k, k, device=dev) * 0.05; b = torch.randn(cout, device=dev)
    oh, ow = (H + s - 1) // s, (W + s - 1) // s
    y = torch.empty(1, cout, oh, ow, device=dev)
    def run():
        _lib.check(_lib.lib().vecvad_fn_conv2d(_lib.ptr(x), x.stride(0), cin, H, W, _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), y.stride(0), cout, k, s, 1, 1,
                                               _lib.ptr(sc), sc.numel(), _lib.cur_stream()), 'conv')
    run(); torch.cuda.synchronize()
    if once: continue
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, w, b, stride=s, padding=(k - 1) // 2), 0.1)
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): run()
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    fl = 2.0 * cin * k * k * cout * oh * ow
    print('%4d->%4d k%d s%d %3dx%3d  %7.1f us  %5.2f TFLOP/s  rel err %.1e' % (cin, cout, k, s, H, W, us, fl / us / 1e6, err), flush=True)
