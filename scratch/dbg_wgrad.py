import sys, torch, numpy as np
sys.path.insert(0, '.')
from vec_vad_b200 import _lib
import torch.nn.functional as F
torch.set_printoptions(precision=4, linewidth=200, sci_mode=False)
def run(x, go, use_tc):
    b, cin, h, wd = x.shape; cout = go.shape[1]
    dw = torch.full((cout, cin, 3, 3), -7.0, device='cuda')
    scratch = torch.empty(9 * cout * cin, device='cuda')
    xn, gn = x.permute(0, 2, 3, 1).contiguous().cuda(), go.permute(0, 2, 3, 1).contiguous().cuda()
    rc = _lib.lib().vecvad_conv3x3_wgrad(_lib.ptr(xn), cin, _lib.ptr(gn), _lib.ptr(dw), _lib.ptr(scratch), b, h, wd, cin, cout, use_tc, _lib.cur_stream())
    _lib.check(rc, 'wgrad'); torch.cuda.synchronize()
    return dw.cpu()
b, cin, cout, h = 2, 32, 32, 32
x = torch.ones(b, cin, h, h) * (torch.arange(cin).view(1, -1, 1, 1) + 1)
go = torch.ones(b, cout, h, h) * (torch.arange(cout).view(1, -1, 1, 1) + 1)
for tc in (0, 1):
    dw = run(x, go, tc)
    print('tc', tc, 'dw[n,k,1,1] block (n<4,k<6):\n', dw[:4, :6, 1, 1], '\n tap corner dw[0,0]:\n', dw[0, 0], 'absmax', dw.abs().max().item(), 'nan', torch.isnan(dw).any().item())
g = torch.Generator().manual_seed(0)
x = torch.randn(b, cin, h, h, generator=g); go = torch.randn(b, cout, h, h, generator=g)
a, c = run(x, go, 0), run(x, go, 1)
print('random: simt', a[0, :4, 1, 1], '\n tc', c[0, :4, 1, 1], 'tc absmax', c.abs().max().item())
