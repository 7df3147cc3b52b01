"""Device probe for the flattened-sequence weight-gradient tiles (wgrad_flat.cu): per-tap errors against autograd."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from vec_vad_b200 import _lib

def run(shape, mode, seed=0):
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, cin, h, wd, generator=g)
    go = torch.randn(b, cout, h, wd, generator=g)
    w = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, padding=1).backward(go.double())
    want = w.grad
    dw = torch.empty(cout, cin, 3, 3, device='cuda')
    scratch = torch.empty(9 * cout * cin, device='cuda')
    xn, gn = x.permute(0, 2, 3, 1).contiguous().cuda(), go.permute(0, 2, 3, 1).contiguous().cuda()
    rc = _lib.lib().vecvad_conv3x3_wgrad(_lib.ptr(xn), cin, _lib.ptr(gn), _lib.ptr(dw), _lib.ptr(scratch), b, h, wd, cin, cout, mode, _lib.cur_stream())
    _lib.check(rc, 'conv3x3_wgrad')
    torch.cuda.synchronize()
    d = (dw.cpu().double() - want).abs()
    scale = want.abs().max().item()
    return d.max().item() / scale, d.amax(dim=(0, 1)) / scale

for shape in [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (130, 32, 32, 32, 32), (2, 32, 32, 32, 64), (2, 64, 64, 32, 32), (5, 16, 16, 32, 32), (1, 32, 32, 32, 32)]:
    try:
        err, per_tap = run(shape, 3)
        print('shape %s err %.3e %s per-tap %s' % (shape, err, 'OK' if err < 3e-3 else 'BAD', ['%.1e' % v for v in per_tap.flatten().tolist()]), flush=True)
    except Exception as e:  # noqa: BLE001
        print('shape', shape, 'EXC', e, flush=True)
        break
