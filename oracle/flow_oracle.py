"""CPU oracle for the FlowNet2 native ops (TEST INFRASTRUCTURE ONLY -- imported by tests/, smoke() and bench.py's
cpu_baseline leg, never by the product package).

Pinned (round 2): the reference ships these ops only as CUDA sources built against ``torch.utils.ffi`` + legacy THC and as sm_30 /
CPython-3.6 binaries, so nothing of it could be executed in round 1.  The kernel SOURCES, however, compile as they are for sm_100a
(``oracle/ref_build/build.sh`` -> ``oracle/_ref/libref_ops.so``, a two-line ``THC.h`` shim for two of the three files);
``tests/golden/make_flow_golden.py`` ran them on a B200 over the cases of ``tests/_flow_cases.py`` and wrote
``tests/golden/flow_ops.npz``, which ``tests/test_flow_reference_golden.py`` holds this file to (CPU) and
``tests/test_flow_ops_gpu.py`` holds the CUDA path to.  ``tests/test_flow_oracle.py`` additionally cross-checks against independent
formulations (torch.nn.functional.unfold / grid_sample / autograd).  This file restates the kernels line by line in numpy (float32
arithmetic where the kernels use float, float64 where they promote to double).

Restated (paths under FlowNet2_src/models/components/ops/):
  correlation_forward ........ correlation/src/correlation_cuda_kernel.cu:10-32 (zero-pad + NHWC repack), :34-106
                               output shape: correlation/src/correlation_cuda.c:25-34
  correlation_backward ....... correlation/src/correlation_cuda_kernel.cu:108-198 (input1), :200-290 (input2)
  resample2d_forward ......... resample2d/src/Resample2d_kernel.cu:20-66
  resample2d_backward ........ resample2d/src/Resample2d_kernel.cu:69-116 (image, trunc quirk), :118-186 (flow)
  channelnorm_forward/backward channelnorm/src/ChannelNorm_kernel.cu:19-51, :54-81
"""
import numpy as np

F32 = np.float32


def correlation_out_shape(h, w, pad_size, kernel_size, max_displacement, stride1, stride2):
    kr = (kernel_size - 1) // 2
    border = kr + max_displacement
    d = (max_displacement // stride2) * 2 + 1
    oh = int(np.ceil(F32(h + 2 * pad_size - 2 * border) / F32(stride1)))
    ow = int(np.ceil(F32(w + 2 * pad_size - 2 * border) / F32(stride1)))
    return d * d, oh, ow


def _pad_nhwc(x, pad):
    """channels_first kernel: NCHW -> zero-padded NHWC (correlation_cuda_kernel.cu:10-32)."""
    n, c, h, w = x.shape
    r = np.zeros((n, h + 2 * pad, w + 2 * pad, c), dtype=F32)
    r[:, pad:pad + h, pad:pad + w, :] = np.transpose(x, (0, 2, 3, 1))
    return r


def correlation_forward(in1, in2, pad_size, kernel_size, max_displacement, stride1, stride2):
    in1, in2 = np.asarray(in1, F32), np.asarray(in2, F32)
    n, c, h, w = in1.shape
    oc, oh, ow = correlation_out_shape(h, w, pad_size, kernel_size, max_displacement, stride1, stride2)
    r1, r2 = _pad_nhwc(in1, pad_size), _pad_nhwc(in2, pad_size)
    kr = (kernel_size - 1) // 2
    dr = max_displacement // stride2
    dsz = 2 * dr + 1
    nelems = F32(kernel_size * kernel_size * c)
    out = np.zeros((n, oc, oh, ow), dtype=F32)
    ys = np.arange(oh) * stride1 + max_displacement + kr        # y1 for every output row   (:53)
    xs = np.arange(ow) * stride1 + max_displacement + kr
    for tj in range(-dr, dr + 1):
        for ti in range(-dr, dr + 1):
            acc = np.zeros((n, oh, ow), dtype=np.float64)
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    a = r1[:, ys[:, None] + j, xs[None, :] + i, :]
                    b = r2[:, ys[:, None] + j + tj * stride2, xs[None, :] + i + ti * stride2, :]
                    acc += np.einsum('nyxc,nyxc->nyx', a, b, dtype=np.float64)
            tc = (tj + dr) * dsz + (ti + dr)
            out[:, tc] = (acc / nelems).astype(F32)
    return out


def _tdiv(a, b):
    """C integer division (truncation toward zero), as the kernels use."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def correlation_backward(in1, in2, grad_out, pad_size, kernel_size, max_displacement, stride1, stride2):
    """Scalar-loop restatement (small cases only)."""
    in1, in2, go = np.asarray(in1, F32), np.asarray(in2, F32), np.asarray(grad_out, F32)
    assert stride1 == 1, 'the reference writes out of bounds for stride1 != 1'
    n, c, h, w = in1.shape
    oc, oh, ow = go.shape[1:]
    r1, r2 = _pad_nhwc(in1, pad_size), _pad_nhwc(in2, pad_size)
    ph, pw = r1.shape[1:3]
    kr = (kernel_size - 1) // 2
    dr = max_displacement // stride2
    dsz = 2 * dr + 1
    nelems = kernel_size * kernel_size * c
    g1, g2 = np.zeros_like(in1), np.zeros_like(in2)
    for yu in range(h):
        for xu in range(w):
            y, x = yu * stride1 + pad_size, xu * stride1 + pad_size
            for which in (1, 2):
                acc = np.zeros((n, c), dtype=np.float64)
                for tc in range(oc):
                    i2 = (tc % dsz - dr) * stride2
                    j2 = (tc // dsz - dr) * stride2
                    sx, sy = (0, 0) if which == 1 else (i2, j2)
                    xmin, ymin = _tdiv(x - kr - max_displacement - sx, stride1), _tdiv(y - kr - max_displacement - sy, stride1)
                    xmax, ymax = _tdiv(x + kr - max_displacement - sx, stride1), _tdiv(y + kr - max_displacement - sy, stride1)
                    if xmax < 0 or ymax < 0 or xmin >= ow or ymin >= oh or xmin > xmax or ymin > ymax:
                        continue
                    xmin, xmax, ymin, ymax = max(0, xmin), min(ow - 1, xmax), max(0, ymin), min(oh - 1, ymax)
                    if which == 1:
                        yy, xx, src = y + j2, x + i2, r2
                    else:
                        yy, xx, src = y - j2, x - i2, r1
                    if not (0 <= yy < ph and 0 <= xx < pw):
                        continue            # the reference would read outside its padded buffer; treated as zero padding
                    val = src[:, yy, xx, :].astype(np.float64)                              # [n, c]
                    s = go[:, tc, ymin:ymax + 1, xmin:xmax + 1].astype(np.float64).sum(axis=(1, 2))   # [n]
                    acc += s[:, None] * val
                (g1 if which == 1 else g2)[:, :, yu, xu] = (acc / nelems).astype(F32)
    return g1, g2


def _bilinear_setup(flow, h_clamp, w_clamp):
    n, _, oh, ow = flow.shape
    xg = np.arange(ow, dtype=F32)[None, None, :]
    yg = np.arange(oh, dtype=F32)[None, :, None]
    xf = (xg + flow[:, 0]).astype(F32)
    yf = (yg + flow[:, 1]).astype(F32)
    fx, fy = np.floor(xf), np.floor(yf)
    xl = np.clip(fx.astype(np.int64), 0, w_clamp - 1)
    xr = np.clip((fx + F32(1)).astype(np.int64), 0, w_clamp - 1)
    yt = np.clip(fy.astype(np.int64), 0, h_clamp - 1)
    yb = np.clip((fy + F32(1)).astype(np.int64), 0, h_clamp - 1)
    return xf, yf, fx, fy, xl, xr, yt, yb


def resample2d_forward(img, flow):
    """kernel_size = 1.  Products in double, accumulated into a float after every term (Resample2d_kernel.cu:57-60)."""
    img, flow = np.asarray(img, F32), np.asarray(flow, F32)
    n, c, ih, iw = img.shape
    oh, ow = flow.shape[2:]
    xf, yf, fx, fy, xl, xr, yt, yb = _bilinear_setup(flow, oh, ow)       # clamped with the OUTPUT size (:48-51)
    alpha, beta = (xf - fx).astype(F32), (yf - fy).astype(F32)
    a64, b64 = alpha.astype(np.float64), beta.astype(np.float64)
    out = np.zeros((n, c, oh, ow), dtype=F32)
    bi = np.arange(n)[:, None, None]
    for ch in range(c):
        im = img[:, ch]
        val = np.zeros((n, oh, ow), dtype=F32)
        for wgt, yy, xx in (((1. - a64) * (1. - b64), yt, xl), (a64 * (1. - b64), yt, xr), ((1. - a64) * b64, yb, xl), (a64 * b64, yb, xr)):
            val = (val.astype(np.float64) + wgt * im[bi, yy, xx].astype(np.float64)).astype(F32)
        out[:, ch] = val
    return out


def resample2d_backward(img, flow, grad_out):
    img, flow, go = np.asarray(img, F32), np.asarray(flow, F32), np.asarray(grad_out, F32)
    n, c, ih, iw = img.shape
    oh, ow = flow.shape[2:]
    # ---- image gradient: scatter-add; weights from xf - int(xf) (truncation), indices clamped with the IMAGE size (:96-105)
    xf, yf, fx, fy, xl, xr, yt, yb = _bilinear_setup(flow, ih, iw)
    alpha = (xf - np.trunc(xf)).astype(F32)
    beta = (yf - np.trunc(yf)).astype(F32)
    g_img = np.zeros((n, c, ih, iw), dtype=np.float64)
    bi = np.broadcast_to(np.arange(n)[:, None, None], xl.shape)
    one = F32(1)
    for ch in range(c):
        g = go[:, ch]
        for wgt, yy, xx in (((one - alpha) * (one - beta), yt, xl), (alpha * (one - beta), yt, xr), ((one - alpha) * beta, yb, xl),
                            (alpha * beta, yb, xr)):
            np.add.at(g_img[:, ch], (bi, yy, xx), (wgt * g).astype(F32))
    # ---- flow gradient (:118-186): indices clamped with the FLOW size
    xf, yf, fx, fy, xl, xr, yt, yb = _bilinear_setup(flow, oh, ow)
    gx = (one - (yf - fy)).astype(F32)     # channel 0 branch uses the y fraction
    gy = (one - (xf - fx)).astype(F32)     # channel 1 branch uses the x fraction
    bi3 = np.arange(n)[:, None, None]
    ox = np.zeros((n, oh, ow), dtype=F32)
    oy = np.zeros((n, oh, ow), dtype=F32)
    for ch in range(c):
        g, im = go[:, ch], img[:, ch]
        tl, tr, bl, br = im[bi3, yt, xl], im[bi3, yt, xr], im[bi3, yb, xl], im[bi3, yb, xr]
        ox = ox + gx * g * tr; ox = ox - gx * g * tl; ox = ox + (one - gx) * g * br; ox = ox - (one - gx) * g * bl
        oy = oy + gy * g * bl; oy = oy - gy * g * tl; oy = oy + (one - gy) * g * br; oy = oy - (one - gy) * g * tr
    return g_img.astype(F32), np.stack([ox, oy], axis=1).astype(F32)


def channelnorm_forward(x):
    x = np.asarray(x, F32)
    acc = np.zeros((x.shape[0],) + x.shape[2:], dtype=F32)
    for c in range(x.shape[1]):
        acc = acc + x[:, c] * x[:, c]
    return np.sqrt(acc)[:, None].astype(F32)


def channelnorm_backward(x, out, grad_out):
    x, out, go = np.asarray(x, F32), np.asarray(out, F32), np.asarray(grad_out, F32)
    return ((go * x).astype(F32).astype(np.float64) / (out.astype(np.float64) + 1e-9)).astype(F32)


def warp_diff_norm(img0, img1, flow):
    """flownet2.py:79-81: resampled = warp(img1, flow); diff = img0 - resampled; norm = channelnorm(diff)."""
    warped = resample2d_forward(img1, flow)
    diff = (np.asarray(img0, F32) - warped).astype(F32)
    return warped, diff, channelnorm_forward(diff)
