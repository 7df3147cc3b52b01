"""The numpy oracle (oracle/flow_oracle.py) held to OUTPUTS OF THE REFERENCE'S OWN KERNELS: tests/golden/flow_ops.npz is
written by tests/golden/make_flow_golden.py, which runs correlation_cuda_kernel.cu / Resample2d_kernel.cu /
ChannelNorm_kernel.cu -- recompiled unmodified for sm_100a (oracle/ref_build/build.sh) -- on the GPU box.  This pins the oracle
for SURVEY.md section 8 rows a12-a15.  CPU-only (no GPU needed: inputs are regenerated from the seeds in tests/_flow_cases.py).

Tolerances: the correlation kernels sum channels in a different order than numpy (fp32 allclose); the
bilinear warp and the channel norm are restated operation by operation (products in double, float accumulation) and must agree
to the last float32 bit or two."""
import os

import numpy as np
import pytest

from oracle import flow_oracle as fo
from tests import _flow_cases as fc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flow_ops.npz')


@pytest.fixture(scope='module')
def gold():
    assert os.path.exists(GOLD), 'tests/golden/flow_ops.npz missing: run tests/golden/make_flow_golden.py on the GPU box'
    with np.load(GOLD, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize('case', fc.CORR_FWD)
def test_oracle_correlation_forward_equals_reference_kernel(case, gold):
    a, b = fc.corr_inputs(case)
    fc.compare(gold, fc.key('corr_fwd', case), fo.correlation_forward(a, b, *case[4:]), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('case', fc.CORR_BWD)
def test_oracle_correlation_backward_equals_reference_kernel(case, gold):
    a, b = fc.corr_inputs(case)
    oc, oh, ow = fo.correlation_out_shape(case[2], case[3], *case[4:])
    go = fc.corr_grad_out(case, (case[0], oc, oh, ow))
    g1, g2 = fo.correlation_backward(a, b, go, *case[4:])
    fc.compare(gold, fc.key('corr_bwd1', case), g1, rtol=2e-4, atol=2e-5)
    fc.compare(gold, fc.key('corr_bwd2', case), g2, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('case', fc.WARP_FWD)
def test_oracle_resample2d_forward_equals_reference_kernel(case, gold):
    img, flow, _ = fc.warp_inputs(case)
    fc.compare(gold, fc.key('warp_fwd', case), fo.resample2d_forward(img, flow), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('case', fc.WARP_BWD)
def test_oracle_resample2d_backward_equals_reference_kernel(case, gold):
    img, flow, go = fc.warp_inputs(case)
    g_img, g_flow = fo.resample2d_backward(img, flow, go)
    fc.compare(gold, fc.key('warp_bwd_img', case), g_img, rtol=1e-4, atol=1e-5)        # atomicAdd order in the reference kernel
    fc.compare(gold, fc.key('warp_bwd_flow', case), g_flow, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('case', fc.NORM)
def test_oracle_channelnorm_equals_reference_kernel(case, gold):
    x, go = fc.norm_inputs(case)
    out = fo.channelnorm_forward(x)
    fc.compare(gold, fc.key('norm_fwd', case), out, rtol=1e-6, atol=1e-7)
    fc.compare(gold, fc.key('norm_bwd', case), fo.channelnorm_backward(x, out, go), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case', fc.WARP_FWD[:2])
def test_oracle_warp_diff_norm_chain_equals_reference_kernels(case, gold):
    img, flow, _ = fc.warp_inputs(case)
    img0 = np.random.RandomState(5000 + sum(int(v) for v in case[:4])).rand(*img.shape).astype(np.float32)
    _, _, norm = fo.warp_diff_norm(img0, img, flow)
    fc.compare(gold, fc.key('chain_norm', case), norm, rtol=1e-5, atol=1e-6)
