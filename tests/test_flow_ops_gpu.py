"""GPU parity of the FlowNet2 ops (libvecvad.so through vec_vad_b200.flow_ops) against oracle/flow_oracle.py on seeded inputs.
Tolerance: fp32 allclose; the summation order over channels differs from the reference's."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from vec_vad_b200 import flow_ops as ops

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# (B, C, H, W, pad, k, md, s1, s2)
CORR = [
    (1, 32, 12, 16, 20, 1, 20, 1, 2),      # FlowNetC parameters (fast path), small map
    (2, 256, 10, 70, 20, 1, 20, 1, 2),     # FlowNetC channels, ragged width (two x tiles, second partial)
    (1, 7, 9, 11, 20, 1, 20, 1, 2),        # channel count not a multiple of the staging chunk
    (2, 5, 9, 11, 4, 1, 4, 1, 2),          # general kernel
    (1, 6, 9, 11, 3, 3, 4, 1, 2),
    (1, 4, 12, 13, 5, 3, 3, 2, 1),
    (1, 3, 8, 9, 0, 1, 2, 1, 2),
]


@pytest.mark.parametrize('case', CORR)
def test_correlation_forward(case):
    b, c, h, w, pad, k, md, s1, s2 = case
    rng = np.random.RandomState(sum(case))
    a, bb = rng.randn(b, c, h, w).astype(np.float32), rng.randn(b, c, h, w).astype(np.float32)
    got = ops.Correlation(pad, k, md, s1, s2, 1)(_t(a), _t(bb)).cpu().numpy()
    want = fo.correlation_forward(a, bb, pad, k, md, s1, s2)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5)


def test_correlation_forward_full_size_properties():
    """BASELINE configs[4] size ([1,256,55,128] maps): size-independent properties instead of the full oracle --
    (a) linearity in the first argument, (b) the zero-displacement channel equals the channel-mean of f1*f2,
    (c) a sampled set of outputs equals the direct dot product."""
    g = torch.Generator().manual_seed(4)
    f1, f1b, f2 = (torch.randn(1, 256, 55, 128, generator=g).cuda() for _ in range(3))
    corr = ops.Correlation(20, 1, 20, 1, 2, 1)
    o1, o1b, o12 = corr(f1, f2), corr(f1b, f2), corr((2 * f1 - 3 * f1b).contiguous(), f2)
    assert tuple(o1.shape) == (1, 441, 55, 128)
    torch.testing.assert_close(o12, 2 * o1 - 3 * o1b, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o1[:, 220], (f1 * f2).mean(1), rtol=1e-4, atol=1e-5)
    rng = np.random.RandomState(0)
    a, b = f1.cpu().double(), f2.cpu().double()
    for _ in range(200):
        tj, ti, y, x = rng.randint(-10, 11), rng.randint(-10, 11), rng.randint(55), rng.randint(128)
        y2, x2 = y + 2 * tj, x + 2 * ti
        want = (a[0, :, y, x] * b[0, :, y2, x2]).sum().item() / 256 if (0 <= y2 < 55 and 0 <= x2 < 128) else 0.0
        assert abs(o1[0, (tj + 10) * 21 + ti + 10, y, x].item() - want) < 1e-4


@pytest.mark.parametrize('case', [(1, 3, 6, 7, 4, 1, 4, 1, 2), (2, 2, 5, 6, 3, 3, 2, 1, 1), (1, 2, 5, 6, 1, 1, 2, 1, 2)])
def test_correlation_backward(case):
    b, c, h, w, pad, k, md, s1, s2 = case
    rng = np.random.RandomState(sum(case))
    a, bb = rng.randn(b, c, h, w).astype(np.float32), rng.randn(b, c, h, w).astype(np.float32)
    ta, tb = _t(a).requires_grad_(), _t(bb).requires_grad_()
    out = ops.Correlation(pad, k, md, s1, s2, 1)(ta, tb)
    go = rng.randn(*out.shape).astype(np.float32)
    out.backward(_t(go))
    g1, g2 = fo.correlation_backward(a, bb, go, pad, k, md, s1, s2)
    np.testing.assert_allclose(ta.grad.cpu().numpy(), g1, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(tb.grad.cpu().numpy(), g2, rtol=2e-4, atol=2e-5)


# the last three run the shared-memory-tiled kernel (width >= 64): ragged tiles, flow far beyond the staged +-12 px window (global
# fallback inside the tiled kernel), one / two / three channels
@pytest.mark.parametrize('shape,amp', [((2, 3, 24, 40), 4.0), ((1, 2, 17, 19), 30.0), ((1, 3, 436, 1024), 4.0), ((2, 3, 70, 150), 30.0),
                                       ((1, 2, 100, 64), 8.0), ((3, 1, 33, 129), 12.0)])
def test_resample2d_forward(shape, amp):
    b, c, h, w = shape
    rng = np.random.RandomState(h)
    img = rng.rand(b, c, h, w).astype(np.float32)
    flow = (rng.randn(b, 2, h, w) * amp).astype(np.float32)      # amp 30 on a 17x19 map: most samples leave the image (border clamp)
    got = ops.Resample2d()(_t(img), _t(flow)).cpu().numpy()
    want = fo.resample2d_forward(img, flow)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)


def test_resample2d_backward():
    rng = np.random.RandomState(8)
    b, c, h, w = 2, 3, 14, 18
    img = rng.rand(b, c, h, w).astype(np.float32)
    flow = (rng.randn(b, 2, h, w) * 3).astype(np.float32)
    go = rng.randn(b, c, h, w).astype(np.float32)
    ti, tf = _t(img).requires_grad_(), _t(flow).requires_grad_()
    ops.Resample2d()(ti, tf).backward(_t(go))
    g_img, g_flow = fo.resample2d_backward(img, flow, go)
    np.testing.assert_allclose(ti.grad.cpu().numpy(), g_img, rtol=1e-4, atol=1e-5)      # atomicAdd order differs
    np.testing.assert_allclose(tf.grad.cpu().numpy(), g_flow, rtol=1e-4, atol=1e-5)


def test_channelnorm_forward_backward():
    rng = np.random.RandomState(9)
    x = rng.randn(2, 3, 20, 33).astype(np.float32)
    t = _t(x).requires_grad_()
    out = ops.ChannelNorm()(t)
    want = fo.channelnorm_forward(x)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    go = rng.randn(2, 1, 20, 33).astype(np.float32)
    out.backward(_t(go))
    np.testing.assert_allclose(t.grad.cpu().numpy(), fo.channelnorm_backward(x, want, go), rtol=1e-5, atol=1e-6)


def test_warp_diff_norm_fused():
    rng = np.random.RandomState(10)
    b, c, h, w = 2, 3, 30, 50
    _fused_case(rng, b, c, h, w, 5)
    _fused_case(rng, 2, 3, 75, 200, 9)            # the tiled kernel (width >= 64), some samples beyond its staged window


def _fused_case(rng, b, c, h, w, amp):
    img0, img1 = rng.rand(b, c, h, w).astype(np.float32), rng.rand(b, c, h, w).astype(np.float32)
    flow = (rng.randn(b, 2, h, w) * amp).astype(np.float32)
    warped, diff, norm = ops.warp_diff_norm(_t(img0), _t(img1), _t(flow))
    w_, d_, n_ = fo.warp_diff_norm(img0, img1, flow)
    np.testing.assert_allclose(warped.cpu().numpy(), w_, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(diff.cpu().numpy(), d_, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(norm.cpu().numpy(), n_, rtol=1e-5, atol=1e-6)
    # and the unfused module chain gives the same thing (flownet2.py:79-81)
    w2 = ops.Resample2d()(_t(img1), _t(flow))
    n2 = ops.ChannelNorm()((_t(img0) - w2).contiguous())
    assert torch.equal(w2, warped)
    np.testing.assert_allclose(n2.cpu().numpy(), norm.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_rejects_cpu_tensors():
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.ChannelNorm()(torch.zeros(1, 2, 3, 3))


# ---------------------------------------------------------------------------------------------------------------------------
# Against OUTPUTS OF THE REFERENCE'S OWN KERNELS (tests/golden/flow_ops.npz, written on a B200 by
# tests/golden/make_flow_golden.py from oracle/_ref/libref_ops.so = the reference .cu files recompiled unmodified).
# ---------------------------------------------------------------------------------------------------------------------------
import os

from tests import _flow_cases as fc

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flow_ops.npz')


@pytest.fixture(scope='module')
def gold():
    assert os.path.exists(_GOLD), 'tests/golden/flow_ops.npz missing'
    with np.load(_GOLD, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize('case', fc.CORR_FWD)
def test_correlation_forward_equals_reference_kernel(case, gold):
    a, b = fc.corr_inputs(case)
    got = ops.Correlation(*case[4:], 1)(_t(a), _t(b)).cpu().numpy()
    fc.compare(gold, fc.key('corr_fwd', case), got, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('case', fc.CORR_BWD)
def test_correlation_backward_equals_reference_kernel(case, gold):
    a, b = fc.corr_inputs(case)
    ta, tb = _t(a).requires_grad_(), _t(b).requires_grad_()
    out = ops.Correlation(*case[4:], 1)(ta, tb)
    out.backward(_t(fc.corr_grad_out(case, tuple(out.shape))))
    fc.compare(gold, fc.key('corr_bwd1', case), ta.grad.cpu().numpy(), rtol=2e-4, atol=2e-5)
    fc.compare(gold, fc.key('corr_bwd2', case), tb.grad.cpu().numpy(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('case', fc.WARP_FWD)
def test_resample2d_forward_equals_reference_kernel(case, gold):
    img, flow, _ = fc.warp_inputs(case)
    fc.compare(gold, fc.key('warp_fwd', case), ops.Resample2d()(_t(img), _t(flow)).cpu().numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('case', fc.WARP_BWD)
def test_resample2d_backward_equals_reference_kernel(case, gold):
    img, flow, go = fc.warp_inputs(case)
    ti, tf = _t(img).requires_grad_(), _t(flow).requires_grad_()
    ops.Resample2d()(ti, tf).backward(_t(go))
    fc.compare(gold, fc.key('warp_bwd_img', case), ti.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    fc.compare(gold, fc.key('warp_bwd_flow', case), tf.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('case', fc.NORM)
def test_channelnorm_equals_reference_kernel(case, gold):
    x, go = fc.norm_inputs(case)
    t = _t(x).requires_grad_()
    out = ops.ChannelNorm()(t)
    fc.compare(gold, fc.key('norm_fwd', case), out.detach().cpu().numpy(), rtol=1e-6, atol=1e-7)
    out.backward(_t(go))
    fc.compare(gold, fc.key('norm_bwd', case), t.grad.cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case', fc.WARP_FWD[:2])
def test_fused_warp_diff_norm_equals_reference_kernel_chain(case, gold):
    img, flow, _ = fc.warp_inputs(case)
    img0 = np.random.RandomState(5000 + sum(int(v) for v in case[:4])).rand(*img.shape).astype(np.float32)
    _, _, norm = ops.warp_diff_norm(_t(img0), _t(img), _t(flow))
    fc.compare(gold, fc.key('chain_norm', case), norm.cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_live_against_recompiled_reference_kernels_at_configs4_size():
    """BASELINE.json configs[4] at full size, element for element, against the reference kernels run live next to ours (the
    recompiled reference library travels to the GPU box inside oracle/_ref; skipped only where it was never built)."""
    from oracle import ref_ops
    if not ref_ops.available():
        pytest.skip('oracle/_ref/libref_ops.so not built (needs /root/reference at build time)')
    g = torch.Generator().manual_seed(12)
    f1, f2 = torch.randn(2, 256, 55, 128, generator=g).cuda(), torch.randn(2, 256, 55, 128, generator=g).cuda()
    got = ops.Correlation(20, 1, 20, 1, 2, 1)(f1, f2)
    want = ref_ops.correlation_forward(f1, f2, 20, 1, 20, 1, 2)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)
    img0, img1 = torch.rand(2, 3, 436, 1024, generator=g).cuda(), torch.rand(2, 3, 436, 1024, generator=g).cuda()
    flow = (torch.randn(2, 2, 436, 1024, generator=g) * 4).cuda()
    w_ref = ref_ops.resample2d_forward(img1, flow)
    torch.testing.assert_close(ops.Resample2d()(img1, flow), w_ref, rtol=1e-6, atol=1e-7)
    n_ref = ref_ops.channelnorm_forward((img0 - w_ref).contiguous())
    warped, diff, norm = ops.warp_diff_norm(img0, img1, flow)
    torch.testing.assert_close(warped, w_ref, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(norm, n_ref, rtol=1e-5, atol=1e-6)
