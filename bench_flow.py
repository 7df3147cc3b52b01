#!/usr/bin/env python
"""Kernel microbench of the FlowNet2 ops at BASELINE.json configs[4]: 1024x436 synthetic frame pairs, 1 GPU.

correlation: conv3 feature maps [B,256,55,128] x2 -> cost volume [B,441,55,128]   (FlowNetC.py:24-30)
warp       : image [B,3,436,1024] + flow [B,2,436,1024] -> warped (+ fused difference and channel norm)
Prints one JSON line per op: device time (CUDA events, L2 flushed between iterations), achieved algorithmic GB/s against
the measured HBM peak (MEASURED_PEAKS.json) and, for the correlation (59 FLOP/B: compute-bound on fp32 FMA), TFLOP/s."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


def run(batch, iters, with_reference=True, device=0):
    """-> one dict per op (ours, then the recompiled reference kernels when oracle/_ref is present and with_reference)."""
    import torch
    from vec_vad_b200 import flow_ops as ops
    pk = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(REPO, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
    dev = torch.device('cuda', device)
    g = torch.Generator().manual_seed(0)
    B = batch
    f1 = torch.randn(B, 256, 55, 128, generator=g).to(dev)
    f2 = torch.randn(B, 256, 55, 128, generator=g).to(dev)
    img0 = torch.rand(B, 3, 436, 1024, generator=g).to(dev)
    img1 = torch.rand(B, 3, 436, 1024, generator=g).to(dev)
    flow = (torch.randn(B, 2, 436, 1024, generator=g) * 4).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    corr = ops.Correlation(20, 1, 20, 1, 2, 1)
    warp = ops.Resample2d()

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            # the 512 MiB memset flushes L2 AND keeps the GPU busy (~150 us) while the host queues the events and the op behind it,
            # so the interval between the events is device time of the op, not host launch latency
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2] * 1e-3
    rows = [
        ('correlation_forward', lambda: corr(f1, f2), B * (2 * 256 * 55 * 128 * 4 + 441 * 55 * 128 * 4), B * 2.0 * 441 * 256 * 55 * 128),
        ('resample2d_forward', lambda: warp(img1, flow), B * (3 + 2 + 3) * 436 * 1024 * 4, 0.0),
        ('warp_diff_norm_fused', lambda: ops.warp_diff_norm(img0, img1, flow), B * (3 + 3 + 2 + 3 + 3 + 1) * 436 * 1024 * 4, 0.0),
    ]
    # the stated bar of SURVEY.md section 2.2: the reference's own kernels, recompiled unmodified for sm_100a (oracle/_ref,
    # test infrastructure -- timed here as the baseline beside ours, never part of the product path)
    have_ref = False
    if with_reference:
        try:
            from oracle import ref_ops
            have_ref = ref_ops.available()
        except ImportError:
            have_ref = False
    if have_ref:
        scratch = torch.empty(ref_ops.lib().ref_correlation_scratch_floats(B, 256, 55, 128, 20), device=dev)
        rows += [
            ('reference_correlation_forward', lambda: ref_ops.correlation_forward(f1, f2, 20, 1, 20, 1, 2, 1, scratch=scratch), rows[0][2], rows[0][3]),
            ('reference_resample2d_forward', lambda: ref_ops.resample2d_forward(img1, flow), rows[1][2], 0.0),
            ('reference_warp_diff_norm_chain', lambda: ref_ops.channelnorm_forward((img0 - ref_ops.resample2d_forward(img1, flow)).contiguous()),
             rows[2][2], 0.0),
        ]
    out = []
    for name, fn, nbytes, flops in rows:
        t = timed(fn)
        line = {'op': name, 'batch': B, 'us': t * 1e6, 'pairs_per_s': B / t, 'algorithmic_bytes': nbytes, 'achieved_gbs': nbytes / t / 1e9,
                'hbm_peak_gbs': pk['hbm_gbs'], 'frac_of_hbm_peak': nbytes / t / 1e9 / pk['hbm_gbs'], 'l2': 'flushed (512 MiB memset) before every timed launch'}
        if flops:
            line['tflops_fp32'] = flops / t / 1e12
        out.append(line)
    return out


def run_flownet2(batch=1, iters=10, height=384, width=512, device=0):
    """The whole FlowNet2 stack (vec_vad_b200/flownet2.py) on the frame size calc_optical_flow.py:49-54 feeds it (512x384), random
    weights: ms per pair and the fp32 FMA rate over the algorithmic conv FLOPs.  Compute-bound on the fp32 pipe (SIMT tiles)."""
    import torch
    from vec_vad_b200 import flownet2 as fn
    dev = torch.device('cuda', device)
    torch.manual_seed(0)
    net = fn.FlowNet2().to(dev).eval()
    x = torch.rand(batch, 3, 2, height, width, device=dev) * 255
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    flops = batch * fn.conv_flops(height, width)
    return {'op': 'flownet2_forward', 'batch': batch, 'frame': '%dx%d' % (width, height), 'ms': t * 1e3, 'pairs_per_s': batch / t,
            'algorithmic_gflop': flops / 1e9, 'tflops_fp32': flops / t / 1e12,
            'note': 'inputs (%d x 3 x 2 x %d x %d fp32) larger than nothing cached between iterations: 162.5 M parameters (650 MB) and '
                    'all activations stream through L2 every pass' % (batch, height, width)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--flownet2', action='store_true', help='time the whole FlowNet2 stack instead of the three ops')
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--no-reference', action='store_true', help='skip the recompiled reference kernels (oracle/_ref)')
    a = ap.parse_args()
    if a.flownet2:
        print(json.dumps(run_flownet2(a.batch, a.iters)))
        return
    for line in run(a.batch, a.iters, with_reference=not a.no_reference):
        print(json.dumps(line))


if __name__ == '__main__':
    main()
