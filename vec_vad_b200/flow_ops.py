"""FlowNet2 native ops on libvecvad.so -- same Python surface as the reference's
``FlowNet2_src/models/components/ops`` package:

  Correlation(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)(in1, in2)   modules/correlation.py:6-27
  Resample2d(kernel_size=1)(img, flow)                                                              modules/resample2d.py:6-14
  ChannelNorm(norm_deg=2)(x)                                                                        modules/channelnorm.py:6-13
  CorrelationFunction / Resample2dFunction / ChannelNormFunction (autograd)                         functions/*.py
plus ``warp_diff_norm(img0, img1, flow)`` -- the fused stage boundary of FlowNet2 (flownet2.py:79-81,93-95,108-115).

Inputs must be contiguous float32 CUDA tensors (the reference asserts contiguity: functions/correlation.py:17-18);
errors surface as RuntimeError (the reference's THError("aborting") also became a Python exception).  No CPU path.
"""
import ctypes as C

import torch
from torch.autograd import Function
from torch.nn.modules.module import Module

from . import _lib


def _chk(*ts):
    _lib.require_cuda(*ts)
    for t in ts:
        if t.dtype != torch.float32:
            raise RuntimeError('vec_vad_b200.flow_ops: float32 tensors only, got %s' % t.dtype)
        assert t.is_contiguous()


class CorrelationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        _chk(input1, input2)
        if input1.shape != input2.shape or input1.dim() != 4:
            raise RuntimeError('Correlation: inputs must be two [B,C,H,W] tensors of the same shape')
        ctx.save_for_backward(input1, input2)
        ctx.args = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        b, c, h, w = input1.shape
        oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
        L = _lib.lib()
        _lib.check(L.vecvad_correlation_out_shape(h, w, pad_size, kernel_size, max_displacement, stride1, stride2, C.byref(oc),
                                                  C.byref(oh), C.byref(ow)), 'correlation_out_shape')
        out = input1.new_empty((b, oc.value, oh.value, ow.value))
        nb = C.c_int64()
        _lib.check(L.vecvad_correlation_workspace_bytes(b, c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2, C.byref(nb)),
                   'correlation_workspace_bytes')
        ws = torch.empty(nb.value // 4, dtype=torch.float32, device=input1.device) if nb.value else None   # caller-owned scratch
        _lib.check(L.vecvad_correlation_forward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(out), b, c, h, w, pad_size, kernel_size,
                                                max_displacement, stride1, stride2, corr_multiply, _lib.ptr(ws), nb.value,
                                                _lib.cur_stream()), 'correlation_forward')
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        b, c, h, w = input1.shape
        g1, g2 = torch.empty_like(input1), torch.empty_like(input2)
        _lib.check(_lib.lib().vecvad_correlation_backward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output), _lib.ptr(g1),
                                                          _lib.ptr(g2), b, c, h, w, *ctx.args, _lib.cur_stream()), 'correlation_backward')
        return (g1, g2) + (None,) * 6


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size, self.kernel_size, self.max_displacement = pad_size, kernel_size, max_displacement
        self.stride1, self.stride2, self.corr_multiply = stride1, stride2, corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement, self.stride1,
                                         self.stride2, self.corr_multiply)


class Resample2dFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1):
        _chk(input1, input2)
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        _, d, ih, iw = input1.shape
        b, _, h, w = input2.shape
        out = input1.new_empty((b, d, h, w))
        _lib.check(_lib.lib().vecvad_resample2d_forward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(out), b, d, ih, iw, h, w, kernel_size,
                                                        _lib.cur_stream()), 'resample2d_forward')
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        _, d, ih, iw = input1.shape
        b, _, h, w = input2.shape
        g1, g2 = torch.empty_like(input1), torch.empty_like(input2)
        _lib.check(_lib.lib().vecvad_resample2d_backward(_lib.ptr(input1), _lib.ptr(input2), _lib.ptr(grad_output), _lib.ptr(g1),
                                                         _lib.ptr(g2), b, d, ih, iw, h, w, ctx.kernel_size, _lib.cur_stream()),
                   'resample2d_backward')
        return g1, g2, None


class Resample2d(Module):
    def __init__(self, kernel_size=1):
        super().__init__()
        self.kernel_size = kernel_size

    def forward(self, input1, input2):
        return Resample2dFunction.apply(input1.contiguous(), input2, self.kernel_size)


class ChannelNormFunction(Function):
    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        _chk(input1)
        b, c, h, w = input1.shape
        out = input1.new_empty((b, 1, h, w))
        _lib.check(_lib.lib().vecvad_channelnorm_forward(_lib.ptr(input1), _lib.ptr(out), b, c, h, w, norm_deg, _lib.cur_stream()),
                   'channelnorm_forward')
        ctx.save_for_backward(input1, out)
        ctx.norm_deg = norm_deg          # (the reference forgets this and its backward raises: functions/channelnorm.py:8-25)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input1, out = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        b, c, h, w = input1.shape
        g = torch.empty_like(input1)
        _lib.check(_lib.lib().vecvad_channelnorm_backward(_lib.ptr(input1), _lib.ptr(out), _lib.ptr(grad_output), _lib.ptr(g), b, c, h, w,
                                                          ctx.norm_deg, _lib.cur_stream()), 'channelnorm_backward')
        return g, None


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)


@torch.no_grad()
def warp_diff_norm(img0, img1, flow, want_diff=True):
    """-> (warped, diff, norm): warped = Resample2d(img1, flow); diff = img0 - warped; norm = ChannelNorm(diff), one kernel."""
    _chk(img0, img1, flow)
    b, c, h, w = img1.shape
    if img0.shape != img1.shape or tuple(flow.shape) != (b, 2, h, w):
        raise RuntimeError('warp_diff_norm: img0/img1 [B,C,H,W] and flow [B,2,H,W] expected')
    warped, norm = torch.empty_like(img1), img1.new_empty((b, 1, h, w))
    diff = torch.empty_like(img1) if want_diff else None
    _lib.check(_lib.lib().vecvad_warp_diff_norm(_lib.ptr(img0), _lib.ptr(img1), _lib.ptr(flow), _lib.ptr(warped), _lib.ptr(diff),
                                                _lib.ptr(norm), b, c, h, w, _lib.cur_stream()), 'warp_diff_norm')
    return warped, diff, norm
