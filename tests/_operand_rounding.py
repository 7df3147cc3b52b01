"""Operand-rounded emulation of the CPU oracle (TEST INFRASTRUCTURE).

The tensor-core tiles round both operands of every 3x3 / transposed convolution -- forward, input-gradient and weight-gradient
contraction -- to a 10-bit mantissa (tf32, or fp16 with a power-of-two loss scale: the same mantissa) and accumulate in fp32.
``rounded_operands()`` patches ``torch.nn.functional.conv2d`` / ``conv_transpose2d`` so that the oracle
(oracle/unet_oracle.py, pinned) does exactly that and NOTHING else: run in fp64, the only difference from the fp64 oracle is
the operand rounding.  The distance emulation <-> fp64 oracle is therefore the error a 10-bit-operand contraction path MUST
show on a given input, and is what tests/test_parity_b128_gpu.py holds the CUDA path's own distance to.

The 1x1 output convolution stays exact (it is fp32 SIMT arithmetic in the CUDA path, unet_kernels.cu k_outconv_fwd).
"""
import contextlib

import torch
import torch.nn.functional as F


def round_mantissa10(t):
    """Round-to-nearest-even to a 10-bit mantissa (tf32 / fp16 precision) without fp16's range limits. fp32 or fp64 in, same dtype out."""
    f = t.detach().to(torch.float32).contiguous()
    bits = f.view(torch.int32)
    lsb = (bits >> 13) & 1
    r = ((bits + 0x0FFF + lsb) & ~0x1FFF).view(torch.float32)
    return r.to(t.dtype)


class _RConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding, round_out):
        xr, wr = round_mantissa10(x), round_mantissa10(w)
        ctx.save_for_backward(xr, wr)
        ctx.padding = padding
        ctx.has_bias = b is not None
        ctx.round_out = round_out
        z = torch.conv2d(xr, wr, b, 1, padding)
        return round_mantissa10(z) if round_out else z

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        gr = round_mantissa10(g)
        gx = torch.nn.grad.conv2d_input(xr.shape, wr, gr, 1, ctx.padding)
        if ctx.round_out:
            gx = round_mantissa10(gx)
        gw = torch.nn.grad.conv2d_weight(xr, wr.shape, gr, 1, ctx.padding)
        gb = g.sum((0, 2, 3)) if ctx.has_bias else None
        return gx, gw, gb, None, None


class _RConvT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding, output_padding, round_out):
        xr, wr = round_mantissa10(x), round_mantissa10(w)
        ctx.save_for_backward(xr, wr)
        ctx.cfg = (stride, padding, output_padding)
        ctx.has_bias = b is not None
        ctx.round_out = round_out
        y = torch.conv_transpose2d(xr, wr, b, stride, padding, output_padding)
        return round_mantissa10(y) if round_out else y

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        stride, padding, output_padding = ctx.cfg
        gr = round_mantissa10(g)
        with torch.enable_grad():
            xx, ww = xr.detach().requires_grad_(True), wr.detach().requires_grad_(True)
            y = torch.conv_transpose2d(xx, ww, None, stride, padding, output_padding)
            gx, gw = torch.autograd.grad(y, (xx, ww), gr)
        if ctx.round_out:
            gx = round_mantissa10(gx)
        gb = g.sum((0, 2, 3)) if ctx.has_bias else None
        return gx, gw, gb, None, None, None, None


@contextlib.contextmanager
def rounded_operands(round_outputs=False):
    """Inside: every conv2d with a kernel larger than 1x1 and every conv_transpose2d rounds its operands (fwd and both gradients).
    round_outputs: the raw conv outputs and the input gradients are rounded as well (the fp16 mode stores them as fp16 in HBM)."""
    conv2d, convT = F.conv2d, F.conv_transpose2d

    def _p(v):
        return (v, v) if isinstance(v, int) else tuple(v)

    def my_conv2d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        if w.shape[-1] == 1 or _p(stride) != (1, 1) or _p(dilation) != (1, 1) or groups != 1:
            return conv2d(x, w, b, stride, padding, dilation, groups)
        return _RConv.apply(x, w, b, _p(padding), round_outputs)

    def my_convT(x, w, b=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
        return _RConvT.apply(x, w, b, _p(stride), _p(padding), _p(output_padding), round_outputs)

    F.conv2d, F.conv_transpose2d = my_conv2d, my_convT
    try:
        yield
    finally:
        F.conv2d, F.conv_transpose2d = conv2d, convT
