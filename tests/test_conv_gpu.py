"""Single-op parity: the 3x3 pad-1 convolution tiles through the C ABI entry point vecvad_conv3x3_forward, against
torch.nn.functional.conv2d in float64 on the CPU (the same op the reference calls: model/unet.py:10,13).
use_tc: 0 = fp32 SIMT tiles, 1 = the tcgen05 tile the engine picks, 3 = flattened-sequence tiles, 4 = pair tiles, +16 = fp16 operands
(kind::f16) instead of tf32.  Tolerances: fp32 tiles 1e-5 of the output range; tf32 / fp16 tiles 2e-3 (10-bit mantissa operands)."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from vec_vad_b200 import _lib

pytestmark = pytest.mark.gpu


def conv3x3(x_nhwc, w, bias, use_tc, want_stats=True):
    b, h, wd, cin = x_nhwc.shape
    cout = w.shape[0]
    out = torch.empty((b, h, wd, cout), device='cuda')
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda') if want_stats else None
    scratch = torch.empty(9 * cout * cin, device='cuda')
    rc = _lib.lib().vecvad_conv3x3_forward(_lib.ptr(x_nhwc), cin, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(stats),
                                           _lib.ptr(scratch), b, h, wd, cin, cout, int(use_tc), _lib.cur_stream())
    _lib.check(rc, 'conv3x3_forward')
    torch.cuda.synchronize()
    return out, stats


SHAPES = [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (5, 16, 16, 32, 64), (3, 8, 8, 64, 128), (5, 4, 4, 128, 256), (9, 4, 4, 256, 256),
          (1, 8, 8, 256, 128), (2, 16, 16, 128, 64)]


@pytest.mark.parametrize('use_tc', [0, 1, 4, 17, 20])
@pytest.mark.parametrize('shape', SHAPES + [(130, 32, 32, 32, 32), (70, 16, 16, 64, 64), (150, 4, 4, 256, 256)])
def test_conv3x3_forward(shape, use_tc):
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(b, cin, h, wd, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).contiguous()
    got, stats = conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), bias.cuda(), use_tc)
    tol = 2e-3 if use_tc else 1e-5
    err = (got.cpu().double() - want).abs().max().item() / want.abs().max().item()
    assert err < tol, err
    # per-channel batch statistics from the epilogue (BatchNorm2d train mode, model/unet.py:11,14)
    s = stats.cpu().numpy()
    np.testing.assert_allclose(s[:cout], want.sum(dim=(0, 1, 2)).numpy(), rtol=0, atol=tol * want.abs().sum(dim=(0, 1, 2)).max().item())
    np.testing.assert_allclose(s[cout:], (want ** 2).sum(dim=(0, 1, 2)).numpy(), rtol=5 * tol)


# flattened-sequence tiles (igemm_flat.cu): any H, W in the tcgen05 tile set with all nine weight tiles resident; ragged batches,
# images whose positions do not fill the last 128-row tile, one and two 32-channel slabs, 32 and 64 output channels
FLAT_SHAPES = [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (2, 32, 32, 32, 64), (130, 32, 32, 32, 32), (2, 64, 64, 32, 32), (5, 16, 16, 32, 64),
               (1, 32, 32, 32, 32), (7, 8, 8, 32, 32), (3, 16, 16, 64, 32), (37, 32, 32, 64, 32)]


@pytest.mark.parametrize('shape', FLAT_SHAPES)
def test_conv3x3_forward_flat(shape):
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape) + 7)
    x = torch.randn(b, cin, h, wd, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    got, stats = conv3x3(xn, w.cuda(), bias.cuda(), 3)
    err = (got.cpu().double() - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-3, err
    s = stats.cpu().numpy()
    np.testing.assert_allclose(s[:cout], want.sum(dim=(0, 1, 2)).numpy(), rtol=0, atol=2e-3 * want.abs().sum(dim=(0, 1, 2)).max().item())
    np.testing.assert_allclose(s[cout:], (want ** 2).sum(dim=(0, 1, 2)).numpy(), rtol=1e-2)
    # same operands, same tf32 rounding, fp32 accumulation in a different order: the two tcgen05 paths agree far below tf32 error
    ref2, _ = conv3x3(xn, w.cuda(), bias.cuda(), 4)
    assert (got - ref2).abs().max().item() <= 2e-5 * want.abs().max().item()


@pytest.mark.parametrize('shape', [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (2, 32, 32, 32, 64), (37, 32, 32, 32, 32), (5, 16, 16, 32, 64)])
def test_conv3x3_forward_flat_fp16_operands(shape):
    """use_tc 16+3: the flattened-sequence tiles with fp16 operands (kind::f16, fp32 accumulation).  fp16 carries tf32's 10-bit
    mantissa, so the bound is the tf32 one."""
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape) + 11)
    x = torch.randn(b, cin, h, wd, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).contiguous()
    got, stats = conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), bias.cuda(), 19)
    err = (got.cpu().double() - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-3, err
    np.testing.assert_allclose(stats.cpu().numpy()[:cout], want.sum(dim=(0, 1, 2)).numpy(), rtol=0,
                               atol=2e-3 * want.abs().sum(dim=(0, 1, 2)).max().item())


def test_tf32_error_is_unbiased():
    """Operands are rounded (not truncated) to tf32 on their way into shared memory: the mean signed error of a
    positive-operand convolution stays far below the truncation bias (~1e-3 relative)."""
    g = torch.Generator().manual_seed(1)
    x = torch.rand(4, 64, 16, 16, generator=g) + 0.5
    w = torch.rand(64, 64, 3, 3, generator=g) + 0.5
    want = F.conv2d(x.double(), w.double(), None, padding=1).permute(0, 2, 3, 1)
    got, _ = conv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), None, 4, want_stats=False)
    rel = ((got.cpu().double() - want) / want)
    assert abs(rel.mean().item()) < 1e-4, rel.mean().item()


# flattened-sequence weight-gradient tiles (wgrad_flat.cu, use_tc 3): all nine taps from one MMA per K-step
WGRAD_FLAT_SHAPES = [(2, 32, 32, 32, 32), (3, 32, 32, 64, 32), (130, 32, 32, 32, 32), (2, 32, 32, 32, 64), (2, 64, 64, 32, 32), (5, 16, 16, 32, 32),
                     (1, 32, 32, 32, 32), (37, 8, 8, 64, 32)]


@pytest.mark.parametrize('use_tc,shape', [(t, s) for t in (0, 1, 2, 17, 18) for s in SHAPES + [(7, 4, 4, 128, 256), (37, 4, 4, 32, 32), (130, 32, 32, 32, 32),
                                                                                           (3, 8, 8, 32, 64)]] + [(t, s) for t in (3, 19) for s in WGRAD_FLAT_SHAPES])
def test_conv3x3_wgrad(shape, use_tc):
    """dW of the convolution = autograd of conv2d (what loss.backward() computes in train.py:401)."""
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = torch.randn(b, cin, h, wd, generator=g)
    go = torch.randn(b, cout, h, wd, generator=g)
    w = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, padding=1).backward(go.double())
    want = w.grad
    dw = torch.empty(cout, cin, 3, 3, device='cuda')
    scratch = torch.empty(9 * cout * cin, device='cuda')
    xn, gn = x.permute(0, 2, 3, 1).contiguous().cuda(), go.permute(0, 2, 3, 1).contiguous().cuda()
    rc = _lib.lib().vecvad_conv3x3_wgrad(_lib.ptr(xn), cin, _lib.ptr(gn), _lib.ptr(dw), _lib.ptr(scratch), b, h, wd, cin, cout, int(use_tc),
                                         _lib.cur_stream())
    _lib.check(rc, 'conv3x3_wgrad')
    err = (dw.cpu().double() - want).abs().max().item() / want.abs().max().item()
    assert err < (3e-3 if use_tc else 2e-5), err


# ---------------------------------------------------------------------------------------------------------------------------
# Input-gradient (flipped taps) and transposed-conv tiles, one op at a time through the C ABI, against float64 autograd of the
# PyTorch ops the reference calls (conv2d: model/unet.py:10,13; ConvTranspose2d(k3, s2, p1, output_padding 1): model/unet.py:54).
# use_tc 1 = the tcgen05 tile the engine picks for the shape (17: with fp16 operands), 0 = fp32 SIMT tiles.  Shapes: the net's own layers (ragged batches).
# ---------------------------------------------------------------------------------------------------------------------------
DGRAD_SHAPES = [(3, 32, 32, 32, 32), (2, 32, 32, 32, 64), (5, 16, 16, 64, 64), (3, 16, 16, 32, 64), (7, 8, 8, 128, 128), (5, 8, 8, 64, 128),
                (9, 4, 4, 256, 256), (37, 4, 4, 128, 256), (130, 32, 32, 32, 32), (3, 8, 8, 256, 128), (2, 16, 16, 128, 64)]


def _tol(use_tc):
    return 3e-3 if use_tc else 2e-5


@pytest.mark.parametrize('use_tc', [0, 1, 17])
@pytest.mark.parametrize('shape', DGRAD_SHAPES)
def test_conv3x3_dgrad(shape, use_tc):
    b, h, wd, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape) + 3)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    go = torch.randn(b, cout, h, wd, generator=g)
    x = torch.zeros(b, cin, h, wd, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), None, padding=1).backward(go.double())
    want = x.grad.permute(0, 2, 3, 1).contiguous()
    gn, wc = go.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda()
    gi = torch.empty((b, h, wd, cin), device='cuda')
    scratch = torch.empty(18 * cout * cin, device='cuda')
    _lib.check(_lib.lib().vecvad_conv3x3_dgrad(_lib.ptr(gn), _lib.ptr(wc), _lib.ptr(gi), _lib.ptr(scratch), b, h, wd, cin, cout, use_tc,
                                               _lib.cur_stream()), 'conv3x3_dgrad')
    torch.cuda.synchronize()
    err = (gi.cpu().double() - want).abs().max().item() / want.abs().max().item()
    assert err < _tol(use_tc), err


CT_SHAPES = [(3, 4, 4, 256, 128), (37, 4, 4, 256, 128), (5, 8, 8, 128, 64), (2, 16, 16, 64, 32), (130, 16, 16, 64, 32), (1, 8, 8, 128, 64)]


def _ct_reference(shape, seed):
    b, h, wd, ci, co = shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, ci, h, wd, generator=g)
    w = torch.randn(ci, co, 3, 3, generator=g) / (1.5 * ci ** 0.5)
    bias = torch.randn(co, generator=g)
    go = torch.randn(b, co, 2 * h, 2 * wd, generator=g)
    xd, wdd = x.double().requires_grad_(), w.double().requires_grad_()
    out = F.conv_transpose2d(xd, wdd, bias.double(), stride=2, padding=1, output_padding=1)
    out.backward(go.double())
    return x, w, bias, go, out.detach(), xd.grad, wdd.grad


@pytest.mark.parametrize('use_tc', [0, 1, 17])
@pytest.mark.parametrize('shape', CT_SHAPES)
def test_conv_transpose_forward_into_concat_half(shape, use_tc):
    """Output lands pixel-shuffled in channels [co, 2co) of a [B,2H,2W,2co] concat buffer (torch.cat([skip, up]), model/unet.py:59);
    the first half must stay untouched."""
    b, h, wd, ci, co = shape
    x, w, bias, go, want, _, _ = _ct_reference(shape, sum(shape) + 5)
    xn, wc, bc = x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda(), bias.cuda()
    cat = torch.full((b, 2 * h, 2 * wd, 2 * co), 7.0, device='cuda')
    scratch = torch.empty(32 * co * ci + co, device='cuda')
    _lib.check(_lib.lib().vecvad_convt3x3s2_forward(_lib.ptr(xn), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(cat), 2 * co, co, _lib.ptr(scratch), b, h,
                                                    wd, ci, co, use_tc, _lib.cur_stream()), 'convt_forward')
    torch.cuda.synchronize()
    got = cat[..., co:].cpu().double()
    want = want.permute(0, 2, 3, 1)
    assert (got - want).abs().max().item() / want.abs().max().item() < _tol(use_tc)
    assert torch.all(cat[..., :co] == 7.0)


@pytest.mark.parametrize('use_tc', [0, 1, 17])
@pytest.mark.parametrize('shape', CT_SHAPES)
def test_conv_transpose_input_and_weight_gradients(shape, use_tc):
    b, h, wd, ci, co = shape
    x, w, bias, go, _, want_gx, want_gw = _ct_reference(shape, sum(shape) + 6)
    xn, wc = x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda()
    # the gradient arrives in the second half of the concat-shaped gradient buffer (net.cu: dCAT), first half holds the skip gradient
    dcat = torch.randn((b, 2 * h, 2 * wd, 2 * co), generator=torch.Generator().manual_seed(1)).cuda()
    dcat[..., co:] = go.permute(0, 2, 3, 1).cuda()
    gi = torch.empty((b, h, wd, ci), device='cuda')
    scratch = torch.empty(48 * co * ci + co, device='cuda')
    L = _lib.lib()
    _lib.check(L.vecvad_convt3x3s2_dgrad(_lib.ptr(dcat), 2 * co, co, _lib.ptr(wc), _lib.ptr(gi), _lib.ptr(scratch), b, h, wd, ci, co, use_tc,
                                         _lib.cur_stream()), 'convt_dgrad')
    torch.cuda.synchronize()
    want = want_gx.permute(0, 2, 3, 1)
    assert (gi.cpu().double() - want).abs().max().item() / want.abs().max().item() < _tol(use_tc)
    dw = torch.empty((ci, co, 3, 3), device='cuda')
    _lib.check(L.vecvad_convt3x3s2_wgrad(_lib.ptr(xn), _lib.ptr(dcat), 2 * co, co, _lib.ptr(dw), _lib.ptr(scratch), b, h, wd, ci, co, use_tc,
                                         _lib.cur_stream()), 'convt_wgrad')
    torch.cuda.synchronize()
    assert (dw.cpu().double() - want_gw).abs().max().item() / want_gw.abs().max().item() < _tol(use_tc)
