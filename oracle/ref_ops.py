"""ctypes binding of oracle/_ref/libref_ops.so: the REFERENCE's own FlowNet2 op kernels (correlation_cuda_kernel.cu,
Resample2d_kernel.cu, ChannelNorm_kernel.cu), compiled unmodified for sm_100a by oracle/ref_build/build.sh.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_flow_golden.py (fixture generation on the GPU box), by the ``-m gpu`` parity
tests (CUDA path vs the reference kernels at full configs[4] size) and by bench_flow.py's ``reference`` rows (the "these sm_30
kernels recompiled" timing bar of SURVEY.md section 2.2).  The product package never imports this module.

Argument meaning follows the reference's Python functions (functions/correlation.py:10-41, functions/resample2d.py:8-21,
functions/channelnorm.py:8-17): contiguous float32 NCHW CUDA tensors in, freshly allocated outputs back.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_ref', 'libref_ops.so')
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError('oracle/_ref/libref_ops.so is missing: run oracle/ref_build/build.sh where /root/reference exists')
        L = C.CDLL(LIB_PATH)
        i, p = C.c_int, C.c_void_p
        ip = C.POINTER(C.c_int)
        L.ref_correlation_out_shape.argtypes = [i] * 7 + [ip, ip, ip]
        L.ref_correlation_out_shape.restype = None
        L.ref_correlation_scratch_floats.argtypes = [i] * 5
        L.ref_correlation_scratch_floats.restype = C.c_longlong
        L.ref_correlation_forward.argtypes = [p, p, p, p] + [i] * 10 + [p]
        L.ref_correlation_backward.argtypes = [p, p, p, p, p, p] + [i] * 10 + [p]
        L.ref_resample2d_forward.argtypes = [p, p, p] + [i] * 7 + [p]
        L.ref_resample2d_backward.argtypes = [p, p, p, p, p] + [i] * 7 + [p]
        L.ref_channelnorm_forward.argtypes = [p, p] + [i] * 5 + [p]
        L.ref_channelnorm_backward.argtypes = [p, p, p, p] + [i] * 5 + [p]
        _lib = L
    return _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ok(rc, what):
    if rc != 1:                       # the reference returns 1 on success (correlation_cuda.c:92)
        raise RuntimeError('reference kernel %s reported failure' % what)


def _chk(*ts):
    for t in ts:
        assert t.is_cuda and t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 4


def correlation_forward(in1, in2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply=1, scratch=None):
    import torch
    _chk(in1, in2)
    b, c, h, w = in1.shape
    oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
    L = lib()
    L.ref_correlation_out_shape(h, w, pad_size, kernel_size, max_displacement, stride1, stride2, C.byref(oc), C.byref(oh), C.byref(ow))
    out = torch.empty((b, oc.value, oh.value, ow.value), device=in1.device)
    if scratch is None:
        scratch = torch.empty(L.ref_correlation_scratch_floats(b, c, h, w, pad_size), device=in1.device)
    _ok(L.ref_correlation_forward(_ptr(in1), _ptr(in2), _ptr(out), _ptr(scratch), b, c, h, w, pad_size, kernel_size, max_displacement,
                                  stride1, stride2, corr_multiply, _stream()), 'Correlation_forward')
    return out


def correlation_backward(in1, in2, grad_out, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply=1):
    import torch
    _chk(in1, in2, grad_out)
    b, c, h, w = in1.shape
    L = lib()
    g1, g2 = torch.empty_like(in1), torch.empty_like(in2)
    scratch = torch.empty(L.ref_correlation_scratch_floats(b, c, h, w, pad_size), device=in1.device)
    _ok(L.ref_correlation_backward(_ptr(in1), _ptr(in2), _ptr(grad_out), _ptr(g1), _ptr(g2), _ptr(scratch), b, c, h, w, pad_size,
                                   kernel_size, max_displacement, stride1, stride2, corr_multiply, _stream()), 'Correlation_backward')
    return g1, g2


def resample2d_forward(img, flow, kernel_size=1, out=None):
    import torch
    _chk(img, flow)
    _, c, ih, iw = img.shape
    b, _, h, w = flow.shape
    if out is None:
        out = torch.empty((b, c, h, w), device=img.device)
    _ok(lib().ref_resample2d_forward(_ptr(img), _ptr(flow), _ptr(out), b, c, ih, iw, h, w, kernel_size, _stream()), 'Resample2d_forward')
    return out


def resample2d_backward(img, flow, grad_out, kernel_size=1):
    import torch
    _chk(img, flow, grad_out)
    _, c, ih, iw = img.shape
    b, _, h, w = flow.shape
    g1, g2 = torch.empty_like(img), torch.empty_like(flow)
    _ok(lib().ref_resample2d_backward(_ptr(img), _ptr(flow), _ptr(grad_out), _ptr(g1), _ptr(g2), b, c, ih, iw, h, w, kernel_size, _stream()),
        'Resample2d_backward')
    return g1, g2


def channelnorm_forward(x, norm_deg=2, out=None):
    import torch
    _chk(x)
    b, c, h, w = x.shape
    if out is None:
        out = torch.empty((b, 1, h, w), device=x.device)
    _ok(lib().ref_channelnorm_forward(_ptr(x), _ptr(out), b, c, h, w, norm_deg, _stream()), 'ChannelNorm_forward')
    return out


def channelnorm_backward(x, out, grad_out, norm_deg=2):
    import torch
    _chk(x, out, grad_out)
    b, c, h, w = x.shape
    g = torch.empty_like(x)
    _ok(lib().ref_channelnorm_backward(_ptr(x), _ptr(out), _ptr(grad_out), _ptr(g), b, c, h, w, norm_deg, _stream()), 'ChannelNorm_backward')
    return g
