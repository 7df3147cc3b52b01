"""small end-to-end pass of every kernel family for compute-sanitizer memcheck"""
import sys, torch
sys.path.insert(0, '.')
from vec_vad_b200 import unet as vu, vad_datasets as vd, flow_ops as ops
kw = dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
for tc in (2, 1, 0):          # fp16-operand tiles, tf32 tiles, fp32 SIMT tiles
    torch.manual_seed(0)
    m = vu.SelfCompleteNet4(use_tensor_cores=tc, **kw).cuda().train()
    m.init_adam()
    g = torch.Generator().manual_seed(1)
    raw = torch.randint(0, 256, (5, 5, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
    fl = torch.randn(5, 1, 32, 32, 2, generator=g).cuda()
    x, xo = vd.cubes_to_device_tensors(raw, fl)
    for _ in range(2):
        l = m.train_step(x, xo)
    m.eval()
    r, o = m.score(x, xo)
    torch.cuda.synchronize()
    print('tc', tc, l.tolist(), r[:2].tolist())
a, b = torch.randn(1, 16, 12, 70, generator=g).cuda(), torch.randn(1, 16, 12, 70, generator=g).cuda()
c = ops.Correlation(20, 1, 20, 1, 2, 1)(a, b)
c2 = ops.Correlation(4, 1, 4, 1, 2, 1)(a, b)
img = torch.rand(1, 3, 20, 30, generator=g).cuda(); f = (torch.randn(1, 2, 20, 30, generator=g) * 5).cuda()
w, d, n = ops.warp_diff_norm(img, img.flip(3).contiguous(), f)
torch.cuda.synchronize()
print('flow ok', c.shape, c2.shape, float(n.sum()))
