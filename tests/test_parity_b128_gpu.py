"""Parity of the path bench.py measures, AT THE BENCHMARK'S BATCH (128 cubes per GPU; BASELINE.json configs[1] and configs[2]).

(a) tensor-core path vs the CPU oracle (oracle/unet_oracle.py, pinned by the reference fixtures) for one train-mode forward +
    backward at B = 128, 5raw1of and 5raw5of: both losses within 1e-4 relative (the north-star bar) and EVERY parameter
    gradient tensor with cosine >= 0.9999 against the oracle's and l2 norm within 1 %.
(b) tensor-core path vs the exact-fp32 SIMT path ON THE DEVICE, same weights, same cubes: per-tensor relative l2 distance under a
    bound derived from the operand rounding, not fitted: unit roundoff u of a 10-bit mantissa (tf32 and fp16 alike) is 2^-11;
    a rounded product carries rms relative error u*sqrt(2/3); a gradient tensor sits behind at most 17 forward + 17 backward
    contractions + its own weight-gradient contraction (35 roundings whose errors add in quadrature) => u*sqrt(2/3)*sqrt(35) =
    2.4e-3; bound = 4 x that = 9.4e-3.
Pre-BN conv biases are excluded everywhere: their gradient is exactly zero here and round-off noise in the reference.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as orc
from tests._util import CONFIGS
from vec_vad_b200 import unet as vu

pytestmark = pytest.mark.gpu
KIND_CLS = {'net4': vu.SelfCompleteNet4, 'full': vu.SelfCompleteNetFull}
B = 128
U = 2.0 ** -11
TC_VS_FP32_BOUND = 4 * U * (2.0 / 3.0) ** 0.5 * 35 ** 0.5
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')


def _is_prebn_bias(name):
    return name.endswith(('conv.0.bias', 'conv.3.bias'))


def _engine_grads(m, x, x_of):
    """One train-mode forward + backward through the engine -> (loss_raw, loss_of, {name: grad})."""
    G = len(m._plan_list)
    sse = torch.empty((G, x.shape[0]), device='cuda')
    m._run_forward(x, x_of, training=True, sse=sse, want_outputs=False)
    m._run_backward(None, None)
    torch.cuda.synchronize()
    flat = m.flat_grads
    grads = {}
    for (name, _), (off, n, shape) in zip(m.named_parameters(), m._param_views):
        grads[name] = flat[off:off + n].view(shape).detach().double().cpu()
    s = sse.double().cpu()
    is_flow = torch.tensor([p[3] for p in m._plan_list], dtype=torch.bool)
    n_raw, n_of = int((~is_flow).sum()), int(is_flow.sum())
    loss_raw = s[~is_flow].sum().item() / (x.shape[0] * 3 * n_raw * 1024)
    loss_of = s[is_flow].sum().item() / (x.shape[0] * 2 * n_of * 1024) if n_of else 0.0
    return loss_raw, loss_of, grads


def _oracle_grads(ref, x, x_of):
    mse = torch.nn.MSELoss()
    ref.train()
    of_out, raw_out, of_tgt, raw_tgt = ref(x, x_of)
    loss_raw = mse(raw_tgt.detach(), raw_out)                  # arguments swapped as in train.py:385
    loss_of = mse(of_tgt.detach(), of_out)
    ref.zero_grad()
    (loss_raw + loss_of).backward()
    return loss_raw.item(), loss_of.item(), {k: p.grad.detach().double() for k, p in ref.named_parameters()}


def _compare(got, want):
    """-> (worst cosine, worst |l2 ratio - 1|, worst relative l2 distance) with the tensors they occur at."""
    rows = []
    for k, w in want.items():
        if _is_prebn_bias(k):
            continue
        g = got[k].reshape(-1)
        w = w.reshape(-1)
        nw, ng = w.norm().item(), g.norm().item()
        cos = float((g @ w).item() / (ng * nw + 1e-300))
        rows.append((k, cos, abs(ng / (nw + 1e-300) - 1.0), (g - w).norm().item() / (nw + 1e-300)))
    return rows


def _report(tag, rows, extra=None):
    worst = {'tag': tag, 'min_cos': min(rows, key=lambda r: r[1])[:2], 'max_l2_dev': max(rows, key=lambda r: r[2])[::2],
             'max_rel_dist': max(rows, key=lambda r: r[3])[::3]}
    worst.update(extra or {})
    try:
        os.makedirs(REPORT, exist_ok=True)
        with open(os.path.join(REPORT, 'parity_b128.jsonl'), 'a') as f:
            f.write(json.dumps(worst) + '\n')
    except OSError:
        pass
    return worst


@pytest.mark.parametrize('prec', [1, 2])                               # 1 = tf32 operands, 2 = fp16 operands
@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2'])          # the configurations (batch comes from this file)
def test_tc_path_matches_oracle_at_benchmark_batch(name, prec):
    kind, kw = CONFIGS[name]
    t_of = kw['tot_of_num']
    torch.manual_seed(17)
    ref = orc.CompletionNetOracle(kind, **kw)
    m = KIND_CLS[kind](use_tensor_cores=prec, **kw)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().train()
    raw_u8, flow = orc.synthetic_cubes(B, t_of=t_of, seed=4321)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    lr_, lo_, want = _oracle_grads(ref, x, x_of)
    gr, go, got = _engine_grads(m, x.cuda(), x_of.cuda())
    rows = _compare(got, want)
    w = _report('tc%d_vs_oracle_%s' % (prec, name), rows, {'loss_raw': [gr, lr_], 'loss_of': [go, lo_]})
    assert abs(gr - lr_) <= 1e-4 * abs(lr_), (gr, lr_)
    assert abs(go - lo_) <= 1e-4 * abs(lo_), (go, lo_)
    assert w['min_cos'][1] >= 0.9999, w
    assert w['max_l2_dev'][1] <= 1e-2, w
    for k in want:                                                  # pre-BN conv biases: exactly zero here
        if _is_prebn_bias(k):
            assert float(got[k].abs().max()) == 0.0


@pytest.mark.parametrize('prec', [1, 2])
@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2'])
def test_tc_path_matches_fp32_simt_path_on_device(name, prec):
    kind, kw = CONFIGS[name]
    raw_u8, flow = orc.synthetic_cubes(B, t_of=kw['tot_of_num'], seed=77)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    x, x_of = x.cuda(), x_of.cuda()
    res = {}
    for tc in (False, prec):
        torch.manual_seed(23)
        m = KIND_CLS[kind](use_tensor_cores=tc, **kw).cuda().train()
        res[tc] = _engine_grads(m, x, x_of)
        del m
    rows = _compare(res[prec][2], res[False][2])
    w = _report('tc%d_vs_simt_%s' % (prec, name), rows, {'loss_raw': [res[prec][0], res[False][0]], 'bound': TC_VS_FP32_BOUND})
    assert abs(res[prec][0] - res[False][0]) <= 1e-4 * abs(res[False][0])
    assert abs(res[prec][1] - res[False][1]) <= 1e-4 * abs(res[False][1])
    assert w['max_rel_dist'][1] <= TC_VS_FP32_BOUND, w
