"""Is the FlowNet2 forward host-bound?  host time to ISSUE one forward vs device time; then the same forward replayed from a CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vec_vad_b200 import flownet2 as fn
torch.manual_seed(0)
net = fn.FlowNet2().cuda().eval()
x = torch.rand(1, 3, 2, 384, 512, device='cuda') * 255
for _ in range(3): net(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): y = net(x)
t_issue = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / 10
print('host issue %.2f ms per forward, wall %.2f ms per forward' % (t_issue * 1e3, t_total * 1e3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    net(x)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    yg = net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g.replay()
e1.record(); torch.cuda.synchronize()
print('graph replay %.2f ms per forward; max |graph - eager| = %.3g' % (e0.elapsed_time(e1) / 10, float((yg - y).abs().max())))
