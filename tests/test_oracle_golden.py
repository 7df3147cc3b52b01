"""Pin the CPU oracle (oracle/unet_oracle.py) against fixtures dumped from the real reference.

These run on CPU (-m "not gpu").  The oracle executes the same ATen ops in the same order as
the reference so the tolerances below only absorb thread-count dependent reduction order.
"""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as orc
from tests._util import CONFIGS, digest, rel_err


def _build(kind, kw, seed):
    torch.manual_seed(seed)
    return orc.CompletionNetOracle(kind, **kw)


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_oracle_matches_reference_fixture(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False)
    kind, kw = CONFIGS[name]
    torch.set_num_threads(1)
    model = _build(kind, kw, int(g['seed_w']))
    # same state_dict keys in the same order, same seeded initial weights
    assert list(model.state_dict().keys()) == [str(k) for k in g['state_keys']]
    _, s0, e0 = digest(model.state_dict().items())
    np.testing.assert_allclose(s0, g['init_stats'], rtol=0, atol=0)
    np.testing.assert_allclose(e0, g['init_samp'], rtol=0, atol=0)
    # dataset adaptor restatement reproduces the reference's input tensors bit-exactly
    x, x_of = orc.cubes_to_tensors(g['raw_u8'], g['flow'])
    assert np.array_equal(x.numpy(), g['x']) and np.array_equal(x_of.numpy(), g['x_of'])
    lam = g['lambda']
    opt = orc.make_adam(model)
    model.train()
    # step 1: outputs, losses, gradients
    import copy
    of_o, raw_o, of_t, raw_t = copy.deepcopy(model)(x, x_of)    # copy: keeps BN running stats of `model` untouched
    assert rel_err(raw_o.detach().numpy(), g['raw_out']) < 1e-6
    assert np.array_equal(raw_t.numpy(), g['raw_tgt'])
    if kw['useFlow']:
        assert rel_err(of_o.detach().numpy(), g['of_out']) < 1e-6
        assert np.array_equal(of_t.numpy(), g['of_tgt'])
    else:
        assert isinstance(of_o, list) and len(of_o) == 0        # reference returns an empty list
    l1 = orc.train_step(model, opt, x, x_of, float(lam[0]), float(lam[1]))
    assert abs(l1[0] - float(g['loss_raw_1'])) <= 1e-6 * abs(float(g['loss_raw_1']))
    assert abs(l1[1] - float(g['loss_of_1'])) <= 1e-6 * max(abs(float(g['loss_of_1'])), 1e-30)
    names, gs, ge = digest([(k, p.grad) for k, p in model.named_parameters()])
    assert [str(n) for n in g['param_names']] == list(names)
    np.testing.assert_allclose(gs[:, 2], g['grad_stats'][:, 2], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(ge, g['grad_samp'], rtol=1e-3, atol=1e-7)
    _, ps, pe = digest(model.state_dict().items())
    np.testing.assert_allclose(ps[:, 2], g['state_stats_1'][:, 2], rtol=1e-5, atol=1e-9)
    # step 2
    l2 = orc.train_step(model, opt, x, x_of, float(lam[0]), float(lam[1]))
    assert abs(l2[0] - float(g['loss_raw_2'])) <= 1e-4 * abs(float(g['loss_raw_2']))
    # eval-mode scoring (train.py:414-427)
    model.eval()
    raw_s, of_s = orc.score_cubes(model, x, x_of)
    np.testing.assert_allclose(raw_s.numpy(), g['score_raw'], rtol=1e-3)
    if kw['useFlow']:
        np.testing.assert_allclose(of_s.numpy(), g['score_of'], rtol=1e-3)


def test_param_counts():
    """12 876 657 params (5raw1of), 21 460 985 (5raw5of)."""
    kind, kw = CONFIGS['net4_flow_b2']
    m = orc.CompletionNetOracle(kind, **kw)
    assert sum(p.numel() for p in m.parameters()) == 12876657
    assert len(list(m.parameters())) == 384
    kind, kw = CONFIGS['full_b2']
    m = orc.CompletionNetOracle(kind, **kw)
    assert sum(p.numel() for p in m.parameters()) == 21460985
    assert len(list(m.parameters())) == 640
