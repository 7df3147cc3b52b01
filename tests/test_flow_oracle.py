"""CPU checks of oracle/flow_oracle.py (parity unpinned by the reference: no runnable reference op, no reference vectors --
see the oracle header).  The restatement is cross-checked against independent formulations built from stock PyTorch ops."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import flow_oracle as fo


def _corr_independent(a, b, pad, k, md, s1, s2):
    """Cost volume via explicit shifts of zero-padded tensors (float64)."""
    a, b = torch.from_numpy(a).double(), torch.from_numpy(b).double()
    n, c, h, w = a.shape
    kr, dr = (k - 1) // 2, md // s2
    big = md + kr + max(pad, 0) + 4
    ap, bp = F.pad(a, [big] * 4), F.pad(b, [big] * 4)
    oc, oh, ow = fo.correlation_out_shape(h, w, pad, k, md, s1, s2)
    out = torch.zeros(n, oc, oh, ow, dtype=torch.float64)
    for y in range(oh):
        for x in range(ow):
            yc, xc = y * s1 + md + kr - pad + big, x * s1 + md + kr - pad + big
            p1 = ap[:, :, yc - kr:yc + kr + 1, xc - kr:xc + kr + 1]
            for tj in range(-dr, dr + 1):
                for ti in range(-dr, dr + 1):
                    p2 = bp[:, :, yc + tj * s2 - kr:yc + tj * s2 + kr + 1, xc + ti * s2 - kr:xc + ti * s2 + kr + 1]
                    out[:, (tj + dr) * (2 * dr + 1) + ti + dr, y, x] = (p1 * p2).sum(dim=(1, 2, 3)) / (k * k * c)
    return out.numpy()


CORR_CASES = [(4, 1, 4, 1, 2), (3, 3, 4, 1, 2), (2, 1, 2, 1, 1), (20, 1, 20, 1, 2), (0, 1, 2, 1, 2), (4, 1, 4, 2, 2), (5, 3, 3, 2, 1)]


@pytest.mark.parametrize('pad,k,md,s1,s2', CORR_CASES)
def test_correlation_forward_vs_shift_formulation(pad, k, md, s1, s2):
    rng = np.random.RandomState(0)
    h, w = (9, 11) if md < 10 else (6, 7)
    a, b = rng.randn(2, 5, h, w).astype(np.float32), rng.randn(2, 5, h, w).astype(np.float32)
    if h + 2 * pad - 2 * (md + (k - 1) // 2) <= 0:
        pytest.skip('empty output')
    got = fo.correlation_forward(a, b, pad, k, md, s1, s2)
    want = _corr_independent(a, b, pad, k, md, s1, s2)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_flownetc_shape():
    """FlowNetC.py:24-30 on a [1,256,48,64] map -> [1,441,48,64]."""
    assert fo.correlation_out_shape(48, 64, 20, 1, 20, 1, 2) == (441, 48, 64)
    assert fo.correlation_out_shape(55, 128, 20, 1, 20, 1, 2) == (441, 55, 128)


@pytest.mark.parametrize('pad,k,md,s2', [(4, 1, 4, 2), (2, 1, 2, 1), (2, 1, 2, 2)])
def test_correlation_backward_is_the_adjoint_when_the_reference_is_consistent(pad, k, md, s2):
    """With kernel_size 1, stride1 1 and pad >= max_displacement (FlowNetC's setting) the reference backward is the true
    gradient of its forward; checked against autograd of the shift formulation.  (For kernel_size > 1 the reference's
    output window [y-kr-md, y+kr-md] is shifted by kr against the true adjoint [y-2kr-md, y-md]: a reference quirk the
    oracle and the CUDA kernel both restate, covered by the oracle-vs-CUDA parity test only.)"""
    rng = np.random.RandomState(1)
    kr = (k - 1) // 2
    if pad < md + kr:
        pytest.skip('reference backward reads outside its padded buffer')
    a, b = rng.randn(1, 3, 6, 7).astype(np.float32), rng.randn(1, 3, 6, 7).astype(np.float32)
    oc, oh, ow = fo.correlation_out_shape(6, 7, pad, k, md, 1, s2)
    go = rng.randn(1, oc, oh, ow).astype(np.float32)
    g1, g2 = fo.correlation_backward(a, b, go, pad, k, md, 1, s2)
    ta, tb = torch.from_numpy(a).double().requires_grad_(), torch.from_numpy(b).double().requires_grad_()
    dr = md // s2
    big = md + kr + pad + 2
    ap, bp = F.pad(ta, [big] * 4), F.pad(tb, [big] * 4)
    tot = 0
    for tj in range(-dr, dr + 1):
        for ti in range(-dr, dr + 1):
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    y0, x0 = md + kr - pad + big + j, md + kr - pad + big + i
                    p1 = ap[:, :, y0:y0 + oh, x0:x0 + ow]
                    p2 = bp[:, :, y0 + tj * s2:y0 + tj * s2 + oh, x0 + ti * s2:x0 + ti * s2 + ow]
                    tot = tot + ((p1 * p2).sum(1) * torch.from_numpy(go[:, (tj + dr) * (2 * dr + 1) + ti + dr]).double()).sum()
    (tot / (k * k * 3)).backward()
    np.testing.assert_allclose(g1, ta.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g2, tb.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_resample2d_forward_vs_grid_sample():
    """Border-clamped bilinear warp == grid_sample(padding_mode='border', align_corners=True) while x+dx stays inside
    [0, W-1] (outside, the reference keeps the unclamped fraction -- exercised by the GPU parity tests instead)."""
    rng = np.random.RandomState(2)
    n, c, h, w = 2, 3, 12, 17
    img = rng.rand(n, c, h, w).astype(np.float32)
    flow = rng.uniform(-3, 3, size=(n, 2, h, w)).astype(np.float32)
    xs, ys = np.arange(w)[None, None, :] + flow[:, 0], np.arange(h)[None, :, None] + flow[:, 1]
    inside = (xs >= 0) & (xs <= w - 1) & (ys >= 0) & (ys <= h - 1)
    got = fo.resample2d_forward(img, flow)
    gx, gy = 2 * xs / (w - 1) - 1, 2 * ys / (h - 1) - 1
    grid = torch.from_numpy(np.stack([gx, gy], -1)).float()
    want = F.grid_sample(torch.from_numpy(img), grid, mode='bilinear', padding_mode='border', align_corners=True).numpy()
    m = np.broadcast_to(inside[:, None], got.shape)
    np.testing.assert_allclose(got[m], want[m], rtol=1e-4, atol=1e-5)
    assert m.mean() > 0.5


def test_resample2d_zero_flow_is_identity_and_integer_shift():
    rng = np.random.RandomState(3)
    img = rng.rand(1, 2, 8, 9).astype(np.float32)
    z = np.zeros((1, 2, 8, 9), np.float32)
    assert np.array_equal(fo.resample2d_forward(img, z), img)
    f = z.copy()
    f[:, 0] = 2.0                                            # sample two pixels to the right, border-clamped
    want = img[:, :, :, np.minimum(np.arange(9) + 2, 8)]
    assert np.array_equal(fo.resample2d_forward(img, f), want)


def test_resample2d_backward_vs_autograd_inside():
    rng = np.random.RandomState(4)
    n, c, h, w = 1, 2, 9, 10
    img = rng.rand(n, c, h, w).astype(np.float32)
    flow = rng.uniform(0.05, 0.95, size=(n, 2, h, w)).astype(np.float32)      # positive -> trunc == floor
    flow[:, 0, :, -1] = -0.5
    flow[:, 1, -1, :] = -0.5                                                    # keep every sample strictly inside
    go = rng.randn(n, c, h, w).astype(np.float32)
    g_img, g_flow = fo.resample2d_backward(img, flow, go)
    ti, tf = torch.from_numpy(img).double().requires_grad_(), torch.from_numpy(flow).double().requires_grad_()
    xs = torch.arange(w).double()[None, None, :] + tf[:, 0]
    ys = torch.arange(h).double()[None, :, None] + tf[:, 1]
    grid = torch.stack([2 * xs / (w - 1) - 1, 2 * ys / (h - 1) - 1], -1)
    out = F.grid_sample(ti, grid, mode='bilinear', padding_mode='border', align_corners=True)
    (out * torch.from_numpy(go).double()).sum().backward()
    np.testing.assert_allclose(g_img, ti.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g_flow, tf.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_channelnorm_vs_torch():
    rng = np.random.RandomState(5)
    x = rng.randn(2, 3, 5, 6).astype(np.float32)
    out = fo.channelnorm_forward(x)
    np.testing.assert_allclose(out, np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True)), rtol=1e-6)
    go = rng.randn(2, 1, 5, 6).astype(np.float32)
    t = torch.from_numpy(x).double().requires_grad_()
    (t.pow(2).sum(1, keepdim=True).sqrt() * torch.from_numpy(go).double()).sum().backward()
    np.testing.assert_allclose(fo.channelnorm_backward(x, out, go), t.grad.numpy(), rtol=1e-5, atol=1e-6)
