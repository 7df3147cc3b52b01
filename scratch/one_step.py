"""one warm-up train step then ONE profiled train step's worth of contraction launches (ncu -c 54 catches the first step only,
so run the kernels twice and let the caller skip): used for the per-step DRAM traffic of the tcgen05 tiles."""
import os, sys, torch
sys.path.insert(0, '.')
from vec_vad_b200 import unet as vu, vad_datasets as vd
kw = dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
torch.manual_seed(0)
m = vu.SelfCompleteNet4(use_tensor_cores=int(os.environ.get('VV_PREC', '2')), **kw).cuda().train()     # 2: fp16 operands (bench default)
m.init_adam()
g = torch.Generator().manual_seed(1)
raw = torch.randint(0, 256, (128, 5, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
fl = torch.randn(128, 1, 32, 32, 2, generator=g).cuda()
x, xo = vd.cubes_to_device_tensors(raw, fl)
for _ in range(int(os.environ.get('VV_STEPS', '2'))):      # the LAST step is the one to read (warm caches, allocations done)
    m.train_step(x, xo)
torch.cuda.synchronize()
