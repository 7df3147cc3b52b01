"""Device probe for the flattened-sequence conv tiles (igemm_flat.cu, use_tc 3): per-shape and per-tap errors against conv2d
(single-tap weights isolate one descriptor start each).  The first version of the kernel could set the descriptor's base-offset
field to the start row's swizzle phase or leave it 0; the run kept as profiles/r01_flat_probe.txt shows that only 0 reads a
SWIZZLE_128B box correctly from a 128-byte-aligned (not 1024-byte-aligned) start, which is what the kernel now does."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from tests.test_conv_gpu import conv3x3  # noqa: E402


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
    got, stats = conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), bias.cuda(), mode)
    d = (got.cpu().double() - want).abs()
    err = d.max().item() / want.abs().max().item()
    s_err = (stats[:cout].cpu() - want.sum(dim=(0, 1, 2))).abs().max().item() / want.abs().sum(dim=(0, 1, 2)).max().item()
    return err, s_err, d


if __name__ == '__main__':
    for mode in (3,):
        for shape in [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (2, 32, 32, 32, 64), (130, 32, 32, 32, 32), (2, 64, 64, 32, 32), (5, 16, 16, 32, 64)]:
            try:
                err, s_err, d = run(shape, mode)
            except Exception as e:  # noqa: BLE001
                print('mode', mode, shape, 'EXC', e)
                continue
            print('mode %d shape %s err %.3e stats_err %.3e %s' % (mode, shape, err, s_err, 'OK' if err < 2e-3 and s_err < 2e-3 else 'BAD'), flush=True)
        shape = (2, 32, 32, 32, 32)
        for tap in range(9):
            err, s_err, d = run(shape, mode, tap)
            bad = (d.amax(dim=3) > 1e-2).nonzero()
            print('  mode %d tap (dy %+d, dx %+d) err %.3e bad pixels %d first %s' % (mode, tap // 3 - 1, tap % 3 - 1, err, bad.shape[0],
                                                                                 bad[:6].tolist()), flush=True)
