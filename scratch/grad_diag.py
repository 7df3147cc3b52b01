"""Per-tensor gradient agreement at the benchmark batch: engine (fp32 SIMT / tensor-core paths) vs the CPU oracle."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as orc
from tests._util import CONFIGS
from tests.test_parity_b128_gpu import _engine_grads, _oracle_grads, _is_prebn_bias, KIND_CLS

name = sys.argv[1] if len(sys.argv) > 1 else 'net4_flow_b2'
modes = [int(a) for a in sys.argv[2:]] or [0, 1]
kind, kw = CONFIGS[name]
torch.manual_seed(17)
ref = orc.CompletionNetOracle(kind, **kw)
raw_u8, flow = orc.synthetic_cubes(128, t_of=kw['tot_of_num'], seed=4321)
x, x_of = orc.cubes_to_tensors(raw_u8, flow)
lr32, lo32, want32 = _oracle_grads(ref, x, x_of)
import copy, time
t0 = time.time()
ref64 = copy.deepcopy(ref).double()
lr_, lo_, want = _oracle_grads(ref64, x.double(), x_of.double())
print('fp64 oracle: %.1f s' % (time.time() - t0))
for tc in [-1] + modes:
  if tc == -1:
    gr, go, got = lr32, lo32, want32
  else:
    m = KIND_CLS[kind](use_tensor_cores=tc, **kw)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().train()
    gr, go, got = _engine_grads(m, x.cuda(), x_of.cuda())
  if True:
    rows = []
    gw, gg = [], []
    for k, w in want.items():
        if _is_prebn_bias(k):
            continue
        g = got[k].reshape(-1); w = w.reshape(-1)
        gw.append(w); gg.append(g)
        rows.append((k, float((g @ w) / (g.norm() * w.norm() + 1e-300)), float(g.norm() / (w.norm() + 1e-300) - 1), float((g - w).norm() / (w.norm() + 1e-300)), float(w.norm()), w.numel()))
    W, G = torch.cat(gw), torch.cat(gg)
    print('== mode', tc, 'loss', gr, lr_, go, lo_, 'global cos', float((G @ W) / (G.norm() * W.norm())), 'global l2 ratio-1', float(G.norm() / W.norm() - 1), 'global rel dist', float((G - W).norm() / W.norm()))
    kinds = {}
    for r in rows:
        k = r[0]
        kk = 'conv_w' if k.endswith(('conv.0.weight', 'conv.3.weight')) else 'bn_w' if k.endswith(('conv.1.weight', 'conv.4.weight')) else 'bn_b' if k.endswith(('conv.1.bias', 'conv.4.bias')) else 'up_w' if k.endswith('up.weight') else 'up_b' if k.endswith('up.bias') else 'out'
        kinds.setdefault(kk, []).append(r)
    for kk, rs in kinds.items():
        print('  %-7s n=%3d  min cos %.6f  max|l2-1| %.4f  max rel dist %.4f   (worst: %s)' % (kk, len(rs), min(r[1] for r in rs), max(abs(r[2]) for r in rs), max(r[3] for r in rs), min(rs, key=lambda r: r[1])[0]))
    rows.sort(key=lambda r: r[1])
    for r in rows[:12]:
        print('   %-40s cos %.6f l2dev %+.4f reldist %.4f |g| %.3e n %d' % r)
