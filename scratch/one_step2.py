"""two train steps (VV_PREC = 1 tf32 / 2 f16 operands, VV_NET = net4 / full): profile with ncu and read the SECOND step's launches."""
import os, sys, torch
sys.path.insert(0, '.')
from vec_vad_b200 import unet as vu, vad_datasets as vd
full = os.environ.get('VV_NET', 'net4') == 'full'
kw = dict(features_root=32, tot_raw_num=5, tot_of_num=5 if full else 1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
torch.manual_seed(0)
cls = vu.SelfCompleteNetFull if full else vu.SelfCompleteNet4
m = cls(use_tensor_cores=int(os.environ.get('VV_PREC', '2')), **kw).cuda().train()
m.init_adam()
g = torch.Generator().manual_seed(1)
raw = torch.randint(0, 256, (128, 5, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
fl = torch.randn(128, 5 if full else 1, 32, 32, 2, generator=g).cuda()
for _ in range(int(os.environ.get('VV_STEPS', '2'))):
    x, xo = vd.cubes_to_device_tensors(raw, fl)
    m.train_step(x, xo)
torch.cuda.synchronize()
