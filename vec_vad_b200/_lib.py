"""ctypes binding of libvecvad.so (C ABI in include/vecvad.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or the
call fails, a RuntimeError is raised.  PyTorch is used by the callers only to own device memory
and streams; every pointer handed to the library is a raw ``tensor.data_ptr()``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvecvad.so')

MAX_UNETS = 10
N_UNITS = 14
N_UPS = 3
ABI_VERSION = 8

# every symbol include/vecvad.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    'vecvad_abi_version', 'vecvad_last_error', 'vecvad_launch_count', 'vecvad_profile_begin', 'vecvad_profile_end',
    'vecvad_correlation_out_shape', 'vecvad_correlation_workspace_bytes', 'vecvad_correlation_forward', 'vecvad_correlation_backward',
    'vecvad_resample2d_forward', 'vecvad_resample2d_backward',
    'vecvad_channelnorm_forward', 'vecvad_channelnorm_backward', 'vecvad_warp_diff_norm',
    'vecvad_net_create', 'vecvad_net_destroy', 'vecvad_net_workspace_bytes', 'vecvad_net_bind',
    'vecvad_net_forward', 'vecvad_net_backward', 'vecvad_net_losses', 'vecvad_net_set_loss_scale', 'vecvad_adam_step',
    'vecvad_net_grad_phase_ranges', 'vecvad_net_grad_phase_wait', 'vecvad_net_defer_join', 'vecvad_adam_step_ranges',
    'vecvad_net_debug_read', 'vecvad_conv3x3_forward', 'vecvad_conv3x3_wgrad', 'vecvad_conv3x3_dgrad',
    'vecvad_convt3x3s2_forward', 'vecvad_convt3x3s2_dgrad', 'vecvad_convt3x3s2_wgrad', 'vecvad_cubes_to_tensors',
    'vecvad_crop_resize',
    'vecvad_fn_conv2d', 'vecvad_fn_deconv4x4s2', 'vecvad_fn_deconv_taps', 'vecvad_fn_conv_plan', 'vecvad_fn_normalize_pair', 'vecvad_fn_upsample4', 'vecvad_fn_scale_copy',
]


class NetConfig(C.Structure):
    """Mirror of ``vecvad_net_config`` (include/vecvad.h)."""
    _fields_ = [
        ('n_unets', C.c_int), ('features_root', C.c_int), ('tot_raw_num', C.c_int), ('patch', C.c_int), ('padding', C.c_int),
        ('param_slot', C.c_int * MAX_UNETS), ('erase_frame', C.c_int * MAX_UNETS), ('out_channels', C.c_int * MAX_UNETS),
        ('target_is_flow', C.c_int * MAX_UNETS), ('target_index', C.c_int * MAX_UNETS), ('out_slot', C.c_int * MAX_UNETS),
        ('slot_param_stride', C.c_int64), ('slot_stat_stride', C.c_int64),
        ('conv_w', C.c_int64 * N_UNITS), ('conv_b', C.c_int64 * N_UNITS), ('bn_w', C.c_int64 * N_UNITS), ('bn_b', C.c_int64 * N_UNITS),
        ('up_w', C.c_int64 * N_UPS), ('up_b', C.c_int64 * N_UPS),
        ('out_w', C.c_int64), ('out_b', C.c_int64),
        ('run_mean', C.c_int64 * N_UNITS), ('run_var', C.c_int64 * N_UNITS),
        ('use_tensor_cores', C.c_int), ('n_raw_total', C.c_int), ('n_of_total', C.c_int),
    ]


_lib = None


def lib():
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('vec_vad_b200: %s is missing -- build it with `python -m vec_vad_b200.build` '
                           '(there is no CPU / PyTorch fallback)' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    p, i, f, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    ip = C.POINTER(C.c_int)
    L.vecvad_abi_version.restype = i
    L.vecvad_last_error.restype = C.c_char_p
    L.vecvad_launch_count.restype = C.c_uint64
    L.vecvad_profile_end.argtypes = [p, p, p, i]
    L.vecvad_correlation_out_shape.argtypes = [i] * 7 + [ip, ip, ip]
    L.vecvad_correlation_workspace_bytes.argtypes = [i] * 9 + [C.POINTER(i64)]
    L.vecvad_correlation_forward.argtypes = [p, p, p] + [i] * 10 + [p, i64, p]
    L.vecvad_correlation_backward.argtypes = [p, p, p, p, p] + [i] * 10 + [p]
    L.vecvad_resample2d_forward.argtypes = [p, p, p] + [i] * 7 + [p]
    L.vecvad_resample2d_backward.argtypes = [p, p, p, p, p] + [i] * 7 + [p]
    L.vecvad_channelnorm_forward.argtypes = [p, p] + [i] * 5 + [p]
    L.vecvad_channelnorm_backward.argtypes = [p, p, p, p] + [i] * 5 + [p]
    L.vecvad_warp_diff_norm.argtypes = [p] * 6 + [i] * 4 + [p]
    L.vecvad_net_create.argtypes = [C.POINTER(NetConfig), C.POINTER(p)]
    L.vecvad_net_destroy.argtypes = [p]
    L.vecvad_net_destroy.restype = None
    L.vecvad_net_workspace_bytes.argtypes = [p, i, C.POINTER(i64)]
    L.vecvad_net_bind.argtypes = [p, p, p, p, p, i64, i]
    L.vecvad_net_forward.argtypes = [p, p, p, i, i, i, p, i, p, i, p, f, f, p]
    L.vecvad_net_backward.argtypes = [p, p, p, p]
    L.vecvad_net_losses.argtypes = [p, p, i, p, p]
    L.vecvad_net_set_loss_scale.argtypes = [p, f]
    L.vecvad_net_grad_phase_ranges.argtypes = [p, C.POINTER(i64), C.POINTER(i64)]
    L.vecvad_net_grad_phase_wait.argtypes = [p, i, p]
    L.vecvad_adam_step.argtypes = [p, p, p, p, i64, f, f, f, f, f, i, f, p]
    L.vecvad_adam_step_ranges.argtypes = [p, p, p, p, i, i64, i64, i64, f, f, f, f, f, i, f, p]
    L.vecvad_net_defer_join.argtypes = [p, i]
    L.vecvad_net_debug_read.argtypes = [p, i, i, p, i64, C.POINTER(i64), p]
    L.vecvad_conv3x3_forward.argtypes = [p, i, p, p, p, p, p, i, i, i, i, i, i, p]
    L.vecvad_conv3x3_wgrad.argtypes = [p, i, p, p, p, i, i, i, i, i, i, p]
    L.vecvad_conv3x3_dgrad.argtypes = [p, p, p, p, i, i, i, i, i, i, p]
    L.vecvad_convt3x3s2_forward.argtypes = [p, p, p, p, i, i, p, i, i, i, i, i, i, p]
    L.vecvad_convt3x3s2_dgrad.argtypes = [p, i, i, p, p, p, i, i, i, i, i, i, p]
    L.vecvad_convt3x3s2_wgrad.argtypes = [p, p, i, i, p, p, i, i, i, i, i, i, p]
    L.vecvad_cubes_to_tensors.argtypes = [p, p, p, p, i, i, i, i, p]
    L.vecvad_crop_resize.argtypes = [p, i, i, i, i, i, i64, i64, i64, i64, p, i, i, p, p]
    L.vecvad_fn_conv2d.argtypes = [p, i64, i, i, i, p, p, p, i64, i, i, i, i, i, p, i64, p]
    L.vecvad_fn_deconv4x4s2.argtypes = [p, i64, i, i, i, p, p, p, i64, i, i, i, p, i64, p]
    L.vecvad_fn_deconv_taps.argtypes = [ip]
    L.vecvad_fn_conv_plan.argtypes = [i, i, i, i, i, i, i, i, i64, ip, ip, ip]
    L.vecvad_fn_normalize_pair.argtypes = [p, p, p, i, i, i, f, p]
    L.vecvad_fn_upsample4.argtypes = [p, i64, i, i, i, p, i64, i, f, i, p]
    L.vecvad_fn_scale_copy.argtypes = [p, i64, p, i64, i64, f, f, i, p]
    if L.vecvad_abi_version() != ABI_VERSION:
        raise RuntimeError('vec_vad_b200: libvecvad.so ABI %d != binding ABI %d -- rebuild' % (L.vecvad_abi_version(), ABI_VERSION))
    _lib = L
    return L


PROFILE_CLASSES = ['conv_dgrad_tcgen05', 'wgrad_tcgen05', 'conv_dgrad_simt', 'wgrad_simt', 'batchnorm', 'other']


def profile_begin():
    check(lib().vecvad_profile_begin(), 'profile_begin')


def profile_end():
    """-> {class: (ms, flops, launches)} accumulated since profile_begin()"""
    n = len(PROFILE_CLASSES)
    ms, fl, la = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
    check(lib().vecvad_profile_end(ms, fl, la, n), 'profile_end')
    return {PROFILE_CLASSES[k]: (ms[k], fl[k], la[k]) for k in range(n)}


def check(rc, what=''):
    if rc != 0:
        msg = lib().vecvad_last_error()
        raise RuntimeError('vec_vad_b200 %s failed (%d): %s' % (what, rc, msg.decode() if msg else '?'))


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('vec_vad_b200 runs on CUDA (sm_100a) only: got a %s tensor -- there is no CPU fallback' % t.device)
