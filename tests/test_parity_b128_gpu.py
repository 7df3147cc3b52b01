"""Parity of the path bench.py measures, AT THE BENCHMARK'S BATCH (128 cubes per GPU; BASELINE.json configs[1] and configs[2]).

What 10-bit operands must cost.  The tensor-core tiles round both operands of every 3x3 / transposed convolution (forward, input
gradient, weight gradient) to a 10-bit mantissa -- tf32, or fp16 with a power-of-two loss scale -- and accumulate in fp32; the
north star mandates those tiles and bounds the MSE at 1e-4 relative.  At B = 128 on random cubes the deep layers' parameter
gradients are sums of heavily cancelling terms: the pinned fp32 CPU oracle itself sits 4e-3 (relative l2) from its own fp64
evaluation, 7e4 times fp32's unit roundoff.  tests/_operand_rounding.py applies exactly the operand rounding (nothing else) to the
fp64 oracle; that emulation sits up to 0.13 from the fp64 oracle on the worst tensors (cosine 0.991) -- the error ANY 10-bit-operand
contraction path has on this input, measured, not assumed.  (profiles/r02_grad_diag_vs_fp64.txt: the CUDA path shows the same.)

(a) tensor-core path vs the oracle, one train-mode forward + backward at B = 128, 5raw1of and 5raw5of, tf32 and fp16 operands:
    * both losses within 1e-4 relative of the fp32 oracle (the north-star bar);
    * vs the fp64 oracle: every parameter gradient tensor no further away (relative l2) than 1.5 x the emulation is (+ 2e-3), and
      the whole gradient (all tensors concatenated) with cosine >= 0.9998;
    * vs the EMULATION: every tensor inside the same ball (1.5 x the emulation's own distance to fp64 + 2e-3).  The realisation of
      the error is not reproducible tensor by tensor -- measured on B200 the CUDA path sits 0.07-0.10 from the emulation where
      both sit 0.13-0.16 from fp64 (profiles/r02_parity_b128.jsonl): the map from a rounding error to these gradients amplifies by
      ~1e5, so fp32-vs-fp64 accumulation order and near-tie roundings decorrelate the two -- hence a magnitude criterion, with
      per-tensor cosine / norm figures REPORTED (gpurun_out/parity_b128.jsonl) rather than bounded at 0.9999 / 1 %, which the
      operand rounding the north star mandates cannot meet on this input (the emulation itself: cosine 0.991 / 0.9875).
    * the exact-fp32 SIMT tiles DO meet cosine >= 0.9999, l2 within 1 % against fp64 (test_fp32_simt_path_...).
(b) tensor-core path vs the exact-fp32 SIMT path ON THE DEVICE, same weights, same cubes: same per-tensor criterion (the SIMT
    path standing in for the oracle: it is itself held to the oracle at 1e-5 on the losses and 1 % on the gradients here).
Pre-BN conv biases are excluded everywhere: their gradient is exactly zero here and round-off noise in the reference.
"""
import copy
import functools
import json
import os

import pytest
import torch

from oracle import unet_oracle as orc
from tests._operand_rounding import rounded_operands
from tests._util import CONFIGS
from vec_vad_b200 import unet as vu

pytestmark = pytest.mark.gpu
KIND_CLS = {'net4': vu.SelfCompleteNet4, 'full': vu.SelfCompleteNetFull}
B = 128
EMU_FACTOR, EMU_FLOOR = 1.5, 2e-3            # distance to fp64 allowed: EMU_FACTOR x the emulation's + EMU_FLOOR
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


@functools.lru_cache(maxsize=None)
def _case(name, seed_w, seed_x):
    """Weights, cubes and the three CPU evaluations of one configuration: fp32 oracle (pinned), fp64 oracle, fp64 + operand rounding
    (tf32 mode: operands only; fp16 mode: raw conv outputs and input gradients too)."""
    kind, kw = CONFIGS[name]
    torch.manual_seed(seed_w)
    ref = orc.CompletionNetOracle(kind, **kw)
    raw_u8, flow = orc.synthetic_cubes(B, t_of=kw['tot_of_num'], seed=seed_x)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    lr32, lo32, _ = _oracle_grads(copy.deepcopy(ref), x, x_of)
    ref64 = copy.deepcopy(ref).double()
    _, _, want64 = _oracle_grads(copy.deepcopy(ref64), x.double(), x_of.double())
    emu = {}
    for prec in (1, 2):
        with rounded_operands(round_outputs=(prec == 2)):
            emu[prec] = _oracle_grads(copy.deepcopy(ref64), x.double(), x_of.double())[2]
    return ref.state_dict(), x, x_of, (lr32, lo32), want64, emu


def _global_cos(got, want):
    g = torch.cat([got[k].reshape(-1) for k in want if not _is_prebn_bias(k)])
    w = torch.cat([want[k].reshape(-1) for k in want if not _is_prebn_bias(k)])
    return float((g @ w) / (g.norm() * w.norm()))


def _held_to_emulation(tag, got, want64, emu, extra):
    """The three gradient criteria of the module docstring; returns the report row."""
    rows_e = _compare(got, emu)                      # vs the emulation: same rounding points
    rows_64 = _compare(got, want64)                  # vs the exact answer ...
    rows_b = {r[0]: r[3] for r in _compare(emu, want64)}   # ... against what the rounding alone costs
    worst_ratio = max(((r[3] - EMU_FLOOR) / rows_b[r[0]], r[0]) for r in rows_64)
    worst_ratio_e = max(((r[3] - EMU_FLOOR) / rows_b[r[0]], r[0]) for r in rows_e)
    w = _report(tag, rows_e, dict(extra, vs_fp64_worst=max(rows_64, key=lambda r: r[3])[::3], emu_vs_fp64_worst=max(rows_b.items(), key=lambda kv: kv[1]),
                                  worst_ratio_to_emu=worst_ratio, worst_ratio_vs_emu=worst_ratio_e, global_cos_fp64=_global_cos(got, want64),
                                  global_cos_emu=_global_cos(got, emu)))
    assert worst_ratio[0] <= EMU_FACTOR, w
    assert worst_ratio_e[0] <= EMU_FACTOR, w
    assert w['global_cos_fp64'] >= 0.9998 and w['global_cos_emu'] >= 0.9998, w
    return w


@pytest.mark.parametrize('prec', [1, 2])                               # 1 = tf32 operands, 2 = fp16 operands
@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2'])          # the configurations (batch comes from this file)
def test_tc_path_matches_oracle_at_benchmark_batch(name, prec):
    kind, kw = CONFIGS[name]
    state, x, x_of, (lr_, lo_), want64, emu = _case(name, 17, 4321)
    m = KIND_CLS[kind](use_tensor_cores=prec, **kw)
    m.load_state_dict(state)
    m = m.cuda().train()
    gr, go, got = _engine_grads(m, x.cuda(), x_of.cuda())
    assert abs(gr - lr_) <= 1e-4 * abs(lr_), (gr, lr_)
    assert abs(go - lo_) <= 1e-4 * abs(lo_), (go, lo_)
    _held_to_emulation('tc%d_vs_oracle_%s' % (prec, name), got, want64, emu[prec], {'loss_raw': [gr, lr_], 'loss_of': [go, lo_]})
    for k in want64:                                                # pre-BN conv biases: exactly zero here
        if _is_prebn_bias(k):
            assert float(got[k].abs().max()) == 0.0


@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2'])
def test_fp32_simt_path_matches_oracle_at_benchmark_batch(name):
    """The exact-fp32 tiles against the fp64 oracle: as close as the fp32 CPU oracle is (4e-3 on the worst tensor), within 2x + 2e-3."""
    kind, kw = CONFIGS[name]
    state, x, x_of, (lr_, lo_), want64, _ = _case(name, 17, 4321)
    m = KIND_CLS[kind](use_tensor_cores=False, **kw)
    m.load_state_dict(state)
    m = m.cuda().train()
    gr, go, got = _engine_grads(m, x.cuda(), x_of.cuda())
    assert abs(gr - lr_) <= 1e-5 * abs(lr_) and abs(go - lo_) <= 1e-5 * abs(lo_)
    rows = _compare(got, want64)
    w = _report('simt_vs_fp64_%s' % name, rows)
    assert w['min_cos'][1] >= 0.9999 and w['max_l2_dev'][1] <= 1e-2 and w['max_rel_dist'][1] <= 1.5e-2, w


@pytest.mark.parametrize('prec', [1, 2])
@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2'])
def test_tc_path_matches_fp32_simt_path_on_device(name, prec):
    """Same weights, same cubes, both paths on the device: the tensor-core path is no further from the fp32 SIMT path than the
    operand-rounding emulation is from the fp64 oracle on this configuration (x 1.5 + 2e-3, per tensor)."""
    kind, kw = CONFIGS[name]
    state, x, x_of, _, want64, emu = _case(name, 17, 4321)
    x, x_of = x.cuda(), x_of.cuda()
    res = {}
    for tc in (False, prec):
        m = KIND_CLS[kind](use_tensor_cores=tc, **kw)
        m.load_state_dict(state)
        m = m.cuda().train()
        res[tc] = _engine_grads(m, x, x_of)
        del m
    rows = _compare(res[prec][2], res[False][2])
    bound = {r[0]: r[3] for r in _compare(emu[prec], want64)}
    worst = max(((r[3] - EMU_FLOOR) / bound[r[0]], r[0]) for r in rows)
    w = _report('tc%d_vs_simt_%s' % (prec, name), rows, {'loss_raw': [res[prec][0], res[False][0]], 'worst_ratio_to_emu': worst,
                                                        'global_cos': _global_cos(res[prec][2], res[False][2])})
    assert abs(res[prec][0] - res[False][0]) <= 1e-4 * abs(res[False][0])
    assert abs(res[prec][1] - res[False][1]) <= 1e-4 * abs(res[False][1])
    assert worst[0] <= EMU_FACTOR, w
    assert w['global_cos'] >= 0.9998, w
