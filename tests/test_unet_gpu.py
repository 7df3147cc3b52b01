"""GPU parity of the completion-UNet engine (libvecvad.so through the module surface) against
(a) the fixtures dumped from the real reference (tests/golden/*.npz) and (b) the CPU oracle on
fresh seeded inputs.  Tolerances are stated per check; the north-star bar is fp32 MSE within 1e-4
relative.  Both contraction paths are covered: fp32 SIMT tiles (tc=False) and tcgen05 tiles (tc=True).
"""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as orc
from tests._util import CONFIGS, digest, rel_err
from vec_vad_b200 import unet as vu

pytestmark = pytest.mark.gpu
KIND_CLS = {'net4': vu.SelfCompleteNet4, 'full': vu.SelfCompleteNetFull, '1raw1of': vu.SelfCompleteNet1raw1of}

# (use_tensor_cores, output rel tol, loss rel tol, grad l2 rel tol)
PATHS = {'simt': (False, 2e-5, 1e-5, 2e-3), 'tc': (True, 5e-3, 1e-4, 6e-2), 'tc16': (2, 5e-3, 1e-4, 6e-2)}     # tc = tf32 operands, tc16 = fp16 operands


def _model(name, g, tc):
    kind, kw = CONFIGS[name]
    torch.manual_seed(int(g['seed_w']))
    return KIND_CLS[kind](use_tensor_cores=tc, **kw).cuda()


@pytest.mark.parametrize('path', sorted(PATHS))
@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_train_forward_backward_matches_reference_fixture(name, path, golden_dir):
    tc, tol_out, tol_loss, tol_grad = PATHS[path]
    g = np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False)
    kind, kw = CONFIGS[name]
    m = _model(name, g, tc)
    x, x_of = torch.from_numpy(g['x']).cuda(), torch.from_numpy(g['x_of']).cuda()
    lam = g['lambda']
    m.train()
    of_o, raw_o, of_t, raw_t = m(x, x_of)
    assert rel_err(raw_o.detach().cpu().numpy(), g['raw_out']) < tol_out
    assert np.array_equal(raw_t.cpu().numpy(), g['raw_tgt'])                    # targets are slices: bit-exact
    mse = torch.nn.MSELoss()
    loss_raw = mse(raw_t.detach(), raw_o)                                        # swapped arguments, train.py:385
    if kw['useFlow']:
        assert rel_err(of_o.detach().cpu().numpy(), g['of_out']) < tol_out
        assert np.array_equal(of_t.cpu().numpy(), g['of_tgt'])
        loss_of = mse(of_t.detach(), of_o)
        loss = float(lam[0]) * loss_raw + float(lam[1]) * loss_of
        assert abs(loss_of.item() - float(g['loss_of_1'])) <= tol_loss * abs(float(g['loss_of_1']))
    else:
        assert isinstance(of_o, list) and len(of_o) == 0
        loss = loss_raw
    assert abs(loss_raw.item() - float(g['loss_raw_1'])) <= tol_loss * abs(float(g['loss_raw_1']))
    loss.backward()
    names, gs, ge = digest([(k, p.grad) for k, p in m.named_parameters()])
    assert [str(n) for n in g['param_names']] == list(names)
    ref_l2 = g['grad_stats'][:, 2]
    # pre-BN conv biases have an exactly-zero gradient (the reference holds round-off noise there): skip those rows
    is_prebn_bias = np.array([n.endswith(('conv.0.bias', 'conv.3.bias')) for n in names])
    big = (~is_prebn_bias) & (ref_l2 > 1e-7)
    np.testing.assert_allclose(gs[big, 2], ref_l2[big], rtol=tol_grad, atol=5e-4 if tc else 2e-6)   # atol: fp32 round-off on near-cancelling sums
    assert np.all(gs[is_prebn_bias, 2] <= 1e-6)
    # complete small gradient tensors, element-wise
    grads = dict((k, p.grad) for k, p in m.named_parameters())
    for key in g.files:
        if key.startswith('grad::') and not key.endswith(('conv.0.bias', 'conv.3.bias')):
            got = grads[key[6:]].cpu().numpy()
            # tf32 path: these per-channel sums cancel heavily at batch 2-4 (|g| ~ 1e-5 from terms ~ 1e-3), hence the wide bound
            assert rel_err(got, g[key]) < (0.3 if tc else 5 * tol_grad), key


@pytest.mark.parametrize('path', sorted(PATHS))
@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2', 'net4_noflow_b4'])
def test_fused_train_step_and_scoring_match_reference_fixture(name, path, golden_dir):
    tc, tol_out, tol_loss, tol_grad = PATHS[path]
    g = np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False)
    kind, kw = CONFIGS[name]
    m = _model(name, g, tc)
    x, x_of = torch.from_numpy(g['x']).cuda(), torch.from_numpy(g['x_of']).cuda()
    lam = g['lambda']
    m.train()
    m.init_adam(lr=1e-3, eps=1e-7)
    l1 = m.train_step(x, x_of, float(lam[0]), float(lam[1])).cpu().numpy().copy()
    assert abs(l1[0] - float(g['loss_raw_1'])) <= tol_loss * abs(float(g['loss_raw_1']))
    if kw['useFlow']:
        assert abs(l1[1] - float(g['loss_of_1'])) <= tol_loss * abs(float(g['loss_of_1']))
    # post-Adam state: first Adam step moves every parameter by ~lr*sign(g): compare l2 norms of every state tensor
    names, ps, pe = digest(m.state_dict().items())
    ref = g['state_stats_1'][:, 2]
    isb = np.array([str(n).endswith(('conv.0.bias', 'conv.3.bias')) for n in names])
    np.testing.assert_allclose(ps[~isb, 2], ref[~isb], rtol=5e-3 if not tc else 2e-2, atol=2e-4)   # atol: |step| < lr where |g| ~ eps
    l2 = m.train_step(x, x_of, float(lam[0]), float(lam[1])).cpu().numpy().copy()
    assert abs(l2[0] - float(g['loss_raw_2'])) <= max(50 * tol_loss, 2e-3) * abs(float(g['loss_raw_2']))


@pytest.mark.parametrize('path', sorted(PATHS))
@pytest.mark.parametrize('name', ['net4_flow_b2', 'full_b2'])
def test_eval_scoring_matches_oracle(name, path, golden_dir):
    """Eval-mode forward with running statistics + per-cube SSE (train.py:414-427) against the oracle with the same state."""
    tc, tol_out, tol_loss, tol_grad = PATHS[path]
    g = np.load(os.path.join(golden_dir, name + '.npz'), allow_pickle=False)
    kind, kw = CONFIGS[name]
    torch.manual_seed(11)
    ref = orc.CompletionNetOracle(kind, **kw)
    with torch.no_grad():                                   # non-trivial running statistics
        for k, b in ref.named_buffers():
            if k.endswith('running_mean'):
                b.normal_(0, 0.1)
            elif k.endswith('running_var'):
                b.uniform_(0.5, 1.5)
    m = KIND_CLS[kind](use_tensor_cores=tc, **kw)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().eval()
    ref.eval()
    x, x_of = torch.from_numpy(g['x']), torch.from_numpy(g['x_of'])
    raw_s, of_s = orc.score_cubes(ref, x, x_of)
    got_raw, got_of = m.score(x.cuda(), x_of.cuda())
    np.testing.assert_allclose(got_raw.cpu().numpy(), raw_s.numpy(), rtol=max(10 * tol_loss, 1e-4))
    np.testing.assert_allclose(got_of.cpu().numpy(), of_s.numpy(), rtol=max(10 * tol_loss, 1e-4))
    with torch.no_grad():
        of_o, raw_o, of_t, raw_t = m(x.cuda(), x_of.cuda())
        rof_o, rraw_o, _, _ = ref(x, x_of)
    assert rel_err(raw_o.cpu().numpy(), rraw_o.numpy()) < tol_out
    assert rel_err(of_o.cpu().numpy(), rof_o.numpy()) < tol_out


@pytest.mark.parametrize('path', sorted(PATHS))
@pytest.mark.parametrize('batch', [3, 16, 37])
def test_against_oracle_fresh_inputs(batch, path):
    """Seeded inputs at ragged / larger batches, three optimiser steps, against the CPU oracle (fp32)."""
    tc, tol_out, tol_loss, tol_grad = PATHS[path]
    kind, kw = CONFIGS['net4_flow_b2']
    torch.manual_seed(5)
    ref = orc.CompletionNetOracle(kind, **kw)
    m = vu.SelfCompleteNet4(use_tensor_cores=tc, **kw)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().train()
    raw_u8, flow = orc.synthetic_cubes(batch, t_of=1, seed=99 + batch)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    opt = orc.make_adam(ref)
    ref.train()
    m.init_adam()
    xc, xoc = x.cuda(), x_of.cuda()
    def check_stats(tol_stat, nsteps):
        # running_mean tracks mean(conv + bias) and the pre-BN conv bias is the one parameter the two implementations
        # treat differently (exactly-zero gradient here, round-off noise pushed through Adam(eps=1e-7) in the reference,
        # SURVEY.md section 7), so the invariant that eval mode uses is compared: running_mean - bias.
        sd_ref, sd = ref.state_dict(), m.state_dict()
        for k in sd_ref:
            if k.endswith('running_mean'):
                kb = k.replace('.1.running_mean', '.0.bias').replace('.4.running_mean', '.3.bias')
                if nsteps == 1:      # both started from the same bias: the first batch mean is directly comparable
                    assert rel_err(sd[k].cpu().numpy(), sd_ref[k].numpy()) < tol_stat, (nsteps, k)
                else:
                    assert rel_err((sd[k] - sd[kb]).cpu().numpy(), (sd_ref[k] - sd_ref[kb]).numpy()) < tol_stat, (nsteps, k)
            if k.endswith('running_var'):
                assert rel_err(sd[k].cpu().numpy(), sd_ref[k].numpy()) < tol_stat, (nsteps, k)
            if k.endswith('num_batches_tracked'):
                assert int(sd[k]) == int(sd_ref[k]) == nsteps

    for step in range(3):
        lr_, lo_ = orc.train_step(ref, opt, x, x_of, 1.0, 1.0)
        got = m.train_step(xc, xoc, 1.0, 1.0).cpu().numpy()
        tol = tol_loss * (1 if step == 0 else 100)          # later steps see Adam(eps=1e-7) amplifying round-off
        assert abs(got[0] - lr_) <= tol * abs(lr_), (step, got, lr_)
        assert abs(got[1] - lo_) <= tol * abs(lo_), (step, got, lo_)
        if step == 0:
            check_stats(1e-4 if not tc else 1e-2, 1)         # statistics of the first batch: no optimiser history involved
    # after three steps: (a) the reference's pre-BN biases have random-walked by up to lr per step (round-off gradients
    # through Adam(eps=1e-7)) and its running_mean has integrated those biases with momentum weights, so running_mean - bias
    # agrees only up to ~3*lr = 3e-3 absolute (values are ~3e-2 at the 4x4 layers, which see batch*16 samples per channel);
    # (b) the split weight-gradient reduction here is summed in a run-dependent order.  Hence a 10 % bound on this check;
    # the tight gates are the step-1 statistics above and the per-step losses.
    check_stats(0.1 if not tc else 0.2, 3)


def test_cubes_to_tensors_bit_exact(golden_dir):
    """Device cube staging == cube_to_train_dataset + ToTensor (vad_datasets.py:130-168), bit for bit."""
    from vec_vad_b200 import vad_datasets as vd
    g = np.load(os.path.join(golden_dir, 'full_b2.npz'), allow_pickle=False)
    x, x_of = vd.cubes_to_device_tensors(torch.from_numpy(g['raw_u8']).cuda(), torch.from_numpy(g['flow']).cuda())
    assert np.array_equal(x.cpu().numpy(), g['x'])
    assert np.array_equal(x_of.cpu().numpy(), g['x_of'])


def test_reference_style_loop_with_torch_adam():
    """The reference's own loop (optim.Adam over model.parameters(), loss.backward(), optimizer.step()) drives the engine."""
    kind, kw = CONFIGS['net4_flow_b2']
    torch.manual_seed(5)
    ref = orc.CompletionNetOracle(kind, **kw)
    m = vu.SelfCompleteNet4(use_tensor_cores=False, **kw)
    m.load_state_dict(ref.state_dict())
    m = m.cuda().train()
    raw_u8, flow = orc.synthetic_cubes(8, t_of=1, seed=3)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    opt_ref = orc.make_adam(ref)
    opt = torch.optim.Adam(m.parameters(), eps=1e-7, weight_decay=0.0)
    mse = torch.nn.MSELoss()
    for step in range(2):
        lr_, lo_ = orc.train_step(ref, opt_ref, x, x_of)
        of_o, raw_o, of_t, raw_t = m(x.cuda(), x_of.cuda())
        loss_raw, loss_of = mse(raw_t.detach(), raw_o), mse(of_t.detach(), of_o)
        opt.zero_grad()
        (loss_raw + loss_of).backward()
        opt.step()
        tol = 1e-5 if step == 0 else 1e-3
        assert abs(loss_raw.item() - lr_) <= tol * lr_ and abs(loss_of.item() - lo_) <= tol * lo_


@pytest.mark.parametrize('tc', [False, True, 2])
def test_side_stream_weight_gradients_equal_single_stream(tc, monkeypatch):
    """The weight-gradient tiles run on a side stream with event-tracked buffer hazards (net.cu).  Same weights, same cubes:
    the flat gradient buffer must equal the single-stream schedule's up to the run-dependent order of the split fp32
    reductions -- partial sums over ~10^5 pixels that cancel to gradients ~10^3 times smaller, measured run-to-run spread
    2e-4 (fp32 tiles) / 1e-3 (tf32 tiles) of the gradient range even on one stream -- repeatedly, at a batch large enough to
    keep both streams busy.  A hazard violation (a tile reading a half-rewritten gradient buffer) shows up at O(1)."""
    kind, kw = CONFIGS['net4_flow_b2']
    raw_u8, flow = orc.synthetic_cubes(64, t_of=1, seed=7)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    x, x_of = x.cuda(), x_of.cuda()
    grads = {}
    for side in ('0', '1'):
        monkeypatch.setenv('VECVAD_WGRAD_STREAM', side)          # read when the engine is created
        torch.manual_seed(3)
        m = vu.SelfCompleteNet4(use_tensor_cores=tc, **kw).cuda().train()
        sse = torch.empty((6, 64), device='cuda')
        runs = []
        for rep in range(4):
            m._run_forward(x, x_of, training=True, sse=sse, want_outputs=False)
            m._run_backward(None, None)
            torch.cuda.synchronize()
            runs.append(m.flat_grads.clone())
        grads[side] = runs
    ref = grads['0'][0]
    scale = ref.abs().max().item()
    for side in ('0', '1'):
        for g_ in grads[side]:
            assert (g_ - ref).abs().max().item() <= (5e-3 if tc else 1e-3) * scale, side


@pytest.mark.parametrize('tc', [False, True, 2])
def test_full_batch_properties(tc):
    """BASELINE.json batch (128 cubes, 5raw1of): properties that need no oracle run.
    (a) the two losses of the fused train step are the means of the per-cube SSE it reports;
    (b) eval-mode scoring is idempotent (bit-identical on a second pass);
    (c) eval-mode scores are per-cube: permuting the batch permutes them, bit for bit (every output row of a tile is
        computed from its own operand row, whichever cubes share the tile);
    (d) a batch made of one cube repeated scores every copy identically."""
    kind, kw = CONFIGS['net4_flow_b2']
    torch.manual_seed(9)
    m = vu.SelfCompleteNet4(use_tensor_cores=tc, **kw).cuda()
    raw_u8, flow = orc.synthetic_cubes(128, t_of=1, seed=21)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    x, x_of = x.cuda(), x_of.cuda()
    m.train()
    m.init_adam()
    sse = torch.empty((6, 128), device='cuda')
    losses = m.train_step(x, x_of, sse=sse).cpu().numpy()
    s = sse.double().cpu().numpy()
    np.testing.assert_allclose(losses[0], s[:5].sum() / (128 * 15 * 1024), rtol=1e-5)
    np.testing.assert_allclose(losses[1], s[5].sum() / (128 * 2 * 1024), rtol=1e-5)
    m.eval()
    r1, o1 = m.score(x, x_of)
    r2, o2 = m.score(x, x_of)
    assert torch.equal(r1, r2) and torch.equal(o1, o2)
    perm = torch.randperm(128, generator=torch.Generator().manual_seed(1)).cuda()
    rp, op_ = m.score(x[perm].contiguous(), x_of[perm].contiguous())
    assert torch.equal(rp, r1[perm]) and torch.equal(op_, o1[perm])
    rr, _ = m.score(x[:1].expand(128, -1, -1, -1).contiguous(), x_of[:1].expand(128, -1, -1, -1).contiguous())
    assert torch.equal(rr, rr[:1].expand(128)) and torch.equal(rr[0], r1[0])


def test_host_cube_feeder_hands_out_the_batches_in_order():
    """HostCubeFeeder: batch i of a cycled list of pinned host batches, converted exactly like cubes_to_device_tensors, one copy ahead."""
    from vec_vad_b200 import vad_datasets as vd
    g = torch.Generator().manual_seed(9)
    host = [(torch.randint(0, 256, (3, 5, 32, 32, 3), generator=g, dtype=torch.uint8).pin_memory(),
             torch.randn((3, 1, 32, 32, 2), generator=g).pin_memory()) for _ in range(3)]
    feeder = vd.HostCubeFeeder(host)
    for i in range(7):                                       # more than one cycle
        x, x_of = feeder.next()
        raw, flow = host[i % 3]
        wx, wo = vd.cubes_to_device_tensors(raw.cuda(), flow.cuda())
        assert torch.equal(x, wx) and torch.equal(x_of, wo)
