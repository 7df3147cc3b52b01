"""Device probe for the fp16-operand variant of the flattened-sequence conv tiles (vecvad_conv3x3_forward, use_tc 5): errors
against conv2d per shape and per tap, next to the tf32 variant (use_tc 3); with VV_TIME=1 also runs both at the bench sizes so
that `ncu --metrics gpu__time_duration.sum -k regex:k_igemm_flat` lists their kernel times."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from vec_vad_b200 import _lib


def conv(x_nhwc, w, bias, mode):
    b, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    out = torch.empty((b, h, wd, cout), device='cuda')
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    nscratch = 9 * cout * cin + (9 * cout * cin) // 2 + 64 + (b * h * wd * cin) // 2 + 64
    scratch = torch.empty(nscratch, device='cuda')
    rc = _lib.lib().vecvad_conv3x3_forward(_lib.ptr(x_nhwc), cin, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(stats),
                                           _lib.ptr(scratch), b, h, wd, cin, cout, int(mode), _lib.cur_stream())
    _lib.check(rc, 'conv3x3_forward')
    torch.cuda.synchronize()
    return out, stats


def run(shape, mode, tap=None, seed=0):
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, cin, h, wd, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    if tap is not None:
        m = torch.zeros(3, 3)
        m[tap // 3, tap % 3] = 1
        w = w * m
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).contiguous()
    got, stats = conv(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), bias.cuda(), mode)
    d = (got.cpu().double() - want).abs()
    return d.max().item() / want.abs().max().item(), d


if __name__ == '__main__':
    if os.environ.get('VV_TC3'):
        # round-2 starting point: the fp16 variant of the pair tiles (use_tc 6, never run on a device in round 1) next to tf32 (4)
        for shape in [(5, 16, 16, 32, 64), (3, 16, 16, 64, 64), (3, 8, 8, 64, 128), (5, 4, 4, 128, 256), (9, 4, 4, 256, 256), (70, 16, 16, 64, 64)]:
            for mode in (4, 6):
                try:
                    err, d = run(shape, mode)
                    print('mode %d shape %s err %.3e %s' % (mode, shape, err, 'OK' if err < 2e-3 else 'BAD'), flush=True)
                except Exception as e:  # noqa: BLE001
                    print('mode', mode, shape, 'EXC', e, flush=True)
        sys.exit(0)
    if os.environ.get('VV_TIME'):
        for cfg in [(768, 32, 32, 32, 32), (768, 32, 32, 64, 32), (768, 32, 32, 32, 64)]:
            for mode in (3, 5):
                b, h, wd, cin, cout = cfg
                x = torch.randn(b, h, wd, cin, device='cuda'); w = torch.randn(cout, cin, 3, 3, device='cuda'); bias = torch.randn(cout, device='cuda')
                for _ in range(3):
                    conv(x, w, bias, mode)
        sys.exit(0)
    for shape in [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (2, 32, 32, 32, 64), (130, 32, 32, 32, 32), (2, 64, 64, 32, 32), (5, 16, 16, 32, 64)]:
        for mode in (3, 5):
            try:
                err, d = run(shape, mode)
                print('mode %d shape %s err %.3e %s' % (mode, shape, err, 'OK' if err < 2e-3 else 'BAD'), flush=True)
            except Exception as e:  # noqa: BLE001
                print('mode', mode, shape, 'EXC', e, flush=True)
    for tap in range(9):
        err, d = run((2, 32, 32, 32, 32), 5, tap)
        bad = (d.amax(dim=3) > 1e-2).nonzero()
        print('  mode 5 tap (dy %+d, dx %+d) err %.3e bad pixels %d first %s' % (tap // 3 - 1, tap % 3 - 1, err, bad.shape[0], bad[:5].tolist()), flush=True)
