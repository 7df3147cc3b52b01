#!/usr/bin/env python
"""Generates tests/golden/flow_ops.npz from the REFERENCE's own CUDA kernels (oracle/_ref/libref_ops.so = the reference's
correlation_cuda_kernel.cu / Resample2d_kernel.cu / ChannelNorm_kernel.cu recompiled unmodified for sm_100a by
oracle/ref_build/build.sh).  The kernels are CUDA, so this runs on the GPU box:

    bash oracle/ref_build/build.sh                                  # here (needs /root/reference)
    gpurun -- 'python tests/golden/make_flow_golden.py gpurun_out/flow_ops.npz'
    cp gpurun_out/flow_ops.npz tests/golden/flow_ops.npz            # then commit

Inputs come from the seeds in tests/_flow_cases.py; the file holds the reference OUTPUTS (full tensors for small cases, a
strided sample + checksums for the configs[4]-sized ones)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import ref_ops  # noqa: E402
from tests import _flow_cases as fc  # noqa: E402


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main(out_path):
    store = {}
    for case in fc.CORR_FWD:
        a, b = fc.corr_inputs(case)
        out = ref_ops.correlation_forward(t(a), t(b), *case[4:])
        fc.pack(store, fc.key('corr_fwd', case), out.cpu().numpy())
    for case in fc.CORR_BWD:
        a, b = fc.corr_inputs(case)
        out = ref_ops.correlation_forward(t(a), t(b), *case[4:])
        go = fc.corr_grad_out(case, tuple(out.shape))
        g1, g2 = ref_ops.correlation_backward(t(a), t(b), t(go), *case[4:])
        fc.pack(store, fc.key('corr_bwd1', case), g1.cpu().numpy())
        fc.pack(store, fc.key('corr_bwd2', case), g2.cpu().numpy())
    for case in fc.WARP_FWD:
        img, flow, _ = fc.warp_inputs(case)
        fc.pack(store, fc.key('warp_fwd', case), ref_ops.resample2d_forward(t(img), t(flow)).cpu().numpy())
    for case in fc.WARP_BWD:
        img, flow, go = fc.warp_inputs(case)
        g1, g2 = ref_ops.resample2d_backward(t(img), t(flow), t(go))
        fc.pack(store, fc.key('warp_bwd_img', case), g1.cpu().numpy())
        fc.pack(store, fc.key('warp_bwd_flow', case), g2.cpu().numpy())
    for case in fc.NORM:
        x, go = fc.norm_inputs(case)
        out = ref_ops.channelnorm_forward(t(x))
        fc.pack(store, fc.key('norm_fwd', case), out.cpu().numpy())
        fc.pack(store, fc.key('norm_bwd', case), ref_ops.channelnorm_backward(t(x), out, t(go)).cpu().numpy())
    # the chain flownet2.py:79-81 runs: warp img1 by the flow, subtract from img0, channel norm
    for case in fc.WARP_FWD[:2]:
        img, flow, go = fc.warp_inputs(case)
        img0 = np.random.RandomState(5000 + sum(int(v) for v in case[:4])).rand(*img.shape).astype(np.float32)
        warped = ref_ops.resample2d_forward(t(img), t(flow))
        diff = (t(img0) - warped).contiguous()
        fc.pack(store, fc.key('chain_norm', case), ref_ops.channelnorm_forward(diff).cpu().numpy())
    torch.cuda.synchronize()
    store['meta::device'] = np.array([ord(c) for c in torch.cuda.get_device_name(0)], dtype=np.uint8)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **store)
    print('wrote', out_path, '%d arrays, %.1f KB' % (len(store), os.path.getsize(out_path) / 1024))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'tests', 'golden', 'flow_ops.npz'))
