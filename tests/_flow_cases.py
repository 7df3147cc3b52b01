"""Seeded cases shared by tests/golden/make_flow_golden.py (which runs the REFERENCE's recompiled kernels on the GPU box and
dumps tests/golden/flow_ops.npz), tests/test_flow_oracle.py (numpy oracle vs those reference outputs, CPU) and
tests/test_flow_ops_gpu.py (CUDA path vs the same outputs).  Inputs are regenerated from the seed (numpy's legacy RandomState
stream is stable across versions and machines), so the fixture holds outputs only; outputs too large to commit are stored as a
strided sample plus float64 checksums."""
import numpy as np

# (B, C, H, W, pad, k, md, s1, s2)
CORR_FWD = [
    (1, 32, 12, 16, 20, 1, 20, 1, 2),      # FlowNetC parameters (FlowNetC.py:24-30), small map
    (2, 256, 10, 70, 20, 1, 20, 1, 2),     # FlowNetC channels, ragged width
    (1, 7, 9, 11, 20, 1, 20, 1, 2),        # odd channel count
    (2, 5, 9, 11, 4, 1, 4, 1, 2),
    (1, 6, 9, 11, 3, 3, 4, 1, 2),
    (1, 4, 12, 13, 5, 3, 3, 2, 1),
    (1, 3, 8, 9, 0, 1, 2, 1, 2),
    (1, 256, 55, 128, 20, 1, 20, 1, 2),    # BASELINE.json configs[4]: conv3 maps of a 1024x436 frame pair
]
# backward cases keep pad >= max_displacement + kernel radius: below that the reference kernels read outside their padded
# repack buffers (correlation_cuda_kernel.cu:150-160), i.e. have no defined output to pin
CORR_BWD = [
    (1, 3, 6, 7, 4, 1, 4, 1, 2),
    (2, 2, 5, 6, 3, 3, 2, 1, 1),
    (1, 8, 6, 7, 20, 1, 20, 1, 2),
]
# (B, C, H, W, flow amplitude in pixels)
WARP_FWD = [(2, 3, 24, 40, 4.0), (1, 2, 17, 19, 30.0), (1, 3, 436, 1024, 4.0)]
WARP_BWD = [(2, 3, 14, 18, 3.0), (1, 3, 33, 47, 6.0)]
NORM = [(2, 3, 20, 33), (1, 2, 17, 64)]

SAMPLE_LIMIT = 40000        # outputs with more elements are stored as a strided sample


def key(prefix, case):
    return prefix + '_' + '_'.join(str(int(v) if float(v).is_integer() else v) for v in case)


def corr_inputs(case):
    b, c, h, w = case[:4]
    rng = np.random.RandomState(1000 + sum(case))
    return rng.randn(b, c, h, w).astype(np.float32), rng.randn(b, c, h, w).astype(np.float32)


def corr_grad_out(case, shape):
    return np.random.RandomState(2000 + sum(case)).randn(*shape).astype(np.float32)


def warp_inputs(case):
    b, c, h, w, amp = case
    rng = np.random.RandomState(3000 + b + c + h + w)
    img = rng.rand(b, c, h, w).astype(np.float32)
    flow = (rng.randn(b, 2, h, w) * amp).astype(np.float32)
    go = rng.randn(b, c, h, w).astype(np.float32)
    return img, flow, go


def norm_inputs(case):
    b, c, h, w = case
    rng = np.random.RandomState(4000 + sum(case))
    return rng.randn(b, c, h, w).astype(np.float32), rng.randn(b, 1, h, w).astype(np.float32)


def pack(store, name, arr):
    """Full tensor when small, else strided sample + checksums."""
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    store[name + '::shape'] = np.array(arr.shape, dtype=np.int64)
    flat = arr.reshape(-1)
    if flat.size <= SAMPLE_LIMIT:
        store[name] = arr
    else:
        step = flat.size // SAMPLE_LIMIT + 1
        store[name + '::step'] = np.array([step], dtype=np.int64)
        store[name + '::sample'] = flat[::step].copy()
        store[name + '::sums'] = np.array([flat.astype(np.float64).sum(), (flat.astype(np.float64) ** 2).sum()])


def compare(store, name, got, rtol, atol):
    """Assert ``got`` equals the stored reference output (full, or sample + checksums)."""
    got = np.ascontiguousarray(got, dtype=np.float32)
    assert tuple(store[name + '::shape']) == got.shape, (name, got.shape, tuple(store[name + '::shape']))
    if name in store:
        np.testing.assert_allclose(got, store[name], rtol=rtol, atol=atol, err_msg=name)
        return
    step = int(store[name + '::step'][0])
    flat = got.reshape(-1)
    np.testing.assert_allclose(flat[::step], store[name + '::sample'], rtol=rtol, atol=atol, err_msg=name)
    s = store[name + '::sums']
    n = flat.size
    # checksums: the element tolerance accumulated in quadrature over n elements (errors are unbiased roundings)
    assert abs(flat.astype(np.float64).sum() - s[0]) <= 10 * (atol + rtol * np.sqrt(s[1] / n)) * np.sqrt(n), name
    np.testing.assert_allclose((flat.astype(np.float64) ** 2).sum(), s[1], rtol=10 * rtol + 1e-6, err_msg=name)
