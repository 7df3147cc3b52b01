"""Per-layer device times of one FlowNet2 forward at 512x384 (synchronising around every layer: diagnosis only)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vec_vad_b200 import flownet2 as fn

rows = []
def wrap(kind):
    orig = getattr(fn._SubNet, kind)
    def timed(self, name, src, dst=None):
        lk, cin, cout, k, s, _ = self.spec[name]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(self, name, src, dst)
        e1.record()
        torch.cuda.synchronize()
        if kind == 'conv':
            fl = 2.0 * src.B * cin * cout * k * k * out.H * out.W
        else:
            fl = 2.0 * src.B * cin * cout * 16 * src.H * src.W
        rows.append((self.kind, name, cin, cout, k, s, src.H, src.W, e0.elapsed_time(e1) * 1e3, fl))
        return out
    setattr(fn._SubNet, kind, timed)
wrap('conv'); wrap('deconv')
torch.manual_seed(0)
net = fn.FlowNet2().cuda().eval()
x = torch.rand(1, 3, 2, 384, 512, device='cuda') * 255
net(x); rows.clear()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); net(x); e1.record(); torch.cuda.synchronize()
tot = sum(r[8] for r in rows)
print('layers %d  sum of layer times %.1f ms  wall (with syncs) %.1f ms' % (len(rows), tot / 1e3, e0.elapsed_time(e1)))
for r in sorted(rows, key=lambda r: -r[8])[:45]:
    print('%-7s %-22s %4d->%4d k%d s%d in %3dx%3d  %8.1f us  %6.2f TFLOP/s' % (r[:8] + (r[8], r[9] / r[8] / 1e6)))
