"""CPU only: how far does a 10-bit-operand contraction path HAVE to be from the fp64 oracle at the benchmark batch?"""
import sys, os, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as orc
from tests._util import CONFIGS
from tests._operand_rounding import rounded_operands
from tests.test_parity_b128_gpu import _oracle_grads, _is_prebn_bias

name = sys.argv[1] if len(sys.argv) > 1 else 'net4_flow_b2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kind, kw = CONFIGS[name]
torch.manual_seed(17)
ref = orc.CompletionNetOracle(kind, **kw)
raw_u8, flow = orc.synthetic_cubes(B, t_of=kw['tot_of_num'], seed=4321)
x, x_of = orc.cubes_to_tensors(raw_u8, flow)
ref64 = copy.deepcopy(ref).double()
t0 = time.time()
lr_, lo_, want = _oracle_grads(ref64, x.double(), x_of.double())
print('fp64 %.1fs' % (time.time() - t0))
t0 = time.time()
with rounded_operands(len(sys.argv) > 3):
    lr_e, lo_e, emu = _oracle_grads(copy.deepcopy(ref64), x.double(), x_of.double())
print('emu %.1fs' % (time.time() - t0), lr_, lr_e, lo_, lo_e)
rows = []
for k, w in want.items():
    if _is_prebn_bias(k): continue
    g = emu[k].reshape(-1); w = w.reshape(-1)
    rows.append((k, float((g @ w) / (g.norm() * w.norm())), float(g.norm() / w.norm() - 1), float((g - w).norm() / w.norm())))
rows.sort(key=lambda r: r[1])
for r in rows[:15]: print('  %-40s cos %.6f l2dev %+.4f reldist %.4f' % r)
print('min cos', rows[0][1], 'max reldist', max(r[3] for r in rows), 'max l2dev', max(abs(r[2]) for r in rows))
