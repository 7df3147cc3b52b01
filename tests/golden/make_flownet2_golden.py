"""Writes tests/golden/flownet2.npz: the REFERENCE's FlowNet2 (FlowNet2_src/models/flownet2.py, imported unmodified from
/root/reference) run on seeded weights and a seeded frame pair.  Run in the build container:
    python tests/golden/make_flownet2_golden.py

The reference's three op packages are cffi bindings to sm_30 / CPython-3.6 binaries that cannot be loaded:
their ``_ext`` modules are replaced here by shims over oracle/flow_oracle.py (itself pinned by tests/golden/flow_ops.npz, written by
the reference's own kernels recompiled for sm_100a).  ``FlowNet2_src.flowlib`` (png / matplotlib, not installed) is not needed by
the model and is shimmed empty.  Everything else -- the module tree, its state_dict keys, the forward graph -- is the reference's.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('VECVAD_REFERENCE', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
from oracle import flow_oracle as fo          # noqa: E402
from oracle import flownet2_oracle as fno     # noqa: E402

OPS = 'FlowNet2_src.models.components.ops.'


def _shim(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _corr_fwd(in1, in2, rbot1, rbot2, out, pad, k, md, s1, s2, mult):
    r = torch.from_numpy(fo.correlation_forward(in1.numpy(), in2.numpy(), pad, k, md, s1, s2))
    out.resize_(r.shape).copy_(r)
    return 1


def _resample_fwd(in1, in2, out, ksize):
    out.copy_(torch.from_numpy(fo.resample2d_forward(in1.numpy(), in2.numpy())))
    return 1


def _norm_fwd(in1, out, deg):
    out.copy_(torch.from_numpy(fo.channelnorm_forward(in1.numpy())))
    return 1


def main():
    _shim('FlowNet2_src.flowlib')
    _shim(OPS + 'correlation._ext', correlation=types.SimpleNamespace(Correlation_forward_cuda=_corr_fwd))
    _shim(OPS + 'resample2d._ext', resample2d=types.SimpleNamespace(Resample2d_cuda_forward=_resample_fwd))
    _shim(OPS + 'channelnorm._ext', channelnorm=types.SimpleNamespace(ChannelNorm_cuda_forward=_norm_fwd))
    from FlowNet2_src.models.flownet2 import FlowNet2      # the reference class
    torch.manual_seed(0)
    net = FlowNet2().eval()
    ref_sd = net.state_dict()
    keys = list(ref_sd.keys())
    shapes = [tuple(v.shape) for v in ref_sd.values()]
    net.load_state_dict(fno.seeded_state(zip(keys, shapes)))
    rng = np.random.default_rng(11)
    H, W = 128, 192
    base = rng.integers(0, 256, (H + 16, W + 16, 3)).astype(np.float32)
    base = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, 2, 0)) / 4.0          # some spatial correlation
    im0, im1 = base[8:8 + H, 8:8 + W], base[5:5 + H, 11:11 + W]                                   # a (3, -3) pixel shift
    frames = np.stack([im0, im1], 0).astype(np.uint8)                                              # [2,H,W,3]
    inputs = torch.from_numpy(frames.astype(np.float32)).permute(3, 0, 1, 2)[None].contiguous()    # [1,3,2,H,W]
    with torch.no_grad():
        out = net(inputs)
        oracle_out, inter = fno.flownet2_forward(net.state_dict(), inputs)
    d = {'frames': frames, 'flow': out.numpy(), 'keys': np.array(keys), 'shapes': np.array([s + (0,) * (4 - len(s)) for s in shapes], dtype=np.int64),
         'ndims': np.array([len(s) for s in shapes], dtype=np.int64)}
    for k in ('flownetc_flow2', 'flownets1_flow2', 'flownets2_flow2', 'flownetsd_flow2'):
        d['oracle_' + k] = inter[k].numpy()
    np.savez_compressed(os.path.join(HERE, 'flownet2.npz'), **d)
    print('params', sum(int(np.prod(s)) for s in shapes), 'keys', len(keys))
    print('flow', out.shape, float(out.abs().mean()), float(out.abs().max()), 'oracle-vs-reference max abs diff', float((out - oracle_out).abs().max()))
    for k in ('flownetc_flow2', 'flownets1_flow2', 'flownets2_flow2', 'flownetsd_flow2'):
        print(k, float(inter[k].abs().mean()))


if __name__ == '__main__':
    main()
