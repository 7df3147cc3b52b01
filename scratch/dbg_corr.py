import sys, torch, numpy as np
sys.path.insert(0, '.')
from vec_vad_b200 import flow_ops as ops
from oracle import flow_oracle as fo
g = torch.Generator().manual_seed(3)
a, b = torch.randn(1, 32, 12, 128, generator=g), torch.randn(1, 32, 12, 128, generator=g)
got = ops.Correlation(20, 1, 20, 1, 2, 1)(a.cuda(), b.cuda())
torch.cuda.synchronize()
want = fo.correlation_forward(a.numpy(), b.numpy(), 20, 1, 20, 1, 2)
print('max err', np.abs(got.cpu().numpy() - want).max())
