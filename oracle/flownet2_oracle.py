"""CPU oracle for the FlowNet2 inference graph (TEST INFRASTRUCTURE ONLY; never imported by vec_vad_b200).

A functional restatement, in plain PyTorch fp32 on the CPU, of the forward pass of the reference's FlowNet2 stack with
``with_bn=False`` (paths under /root/reference/FlowNet2_src/models):
  * FlowNet2.forward ........ flownet2.py:65-149
  * FlowNetC.forward ........ components/FlowNetC.py:75-132
  * FlowNetS.forward ........ components/FlowNetS.py:59-96
  * FlowNetSD.forward ....... components/FlowNetSD.py:55-103
  * FlowNetFusion.forward ... components/FlowNetFusion.py:45-64
  * conv / deconv / predict_flow building blocks ... components/misc.py:6-45
It consumes a ``state_dict`` with the reference's own keys, so the same weights drive the reference, this oracle and the CUDA path.
The three native ops come from oracle/flow_oracle.py (pinned by tests/golden/flow_ops.npz, written by the reference's own kernels).

Pinning: tests/golden/make_flownet2_golden.py imports the reference's FlowNet2 class itself (its cffi op packages replaced by
shims over oracle/flow_oracle.py -- the reference's op binaries cannot be loaded), runs it on seeded weights
and inputs and stores the result in tests/golden/flownet2.npz; tests/test_flownet2.py holds this file to that fixture on CPU.
nn.Upsample(scale_factor=4, mode='bilinear') is evaluated with align_corners=False, PyTorch's behaviour from 0.4 on (README pins
PyTorch 1.1.0).
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import flow_oracle as fo


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def correlation(a, b):
    return _t(fo.correlation_forward(a.numpy(), b.numpy(), 20, 1, 20, 1, 2))          # FlowNetC.py:24-30


def resample(img, flow):
    return _t(fo.resample2d_forward(np.ascontiguousarray(img.numpy()), np.ascontiguousarray(flow.numpy())))


def channelnorm(x):
    return _t(fo.channelnorm_forward(np.ascontiguousarray(x.numpy())))


class _Net:
    def __init__(self, sd, prefix):
        self.sd, self.p = sd, prefix

    def _wb(self, name, wrapped):
        k = self.p + name + ('.0' if wrapped else '')
        return self.sd[k + '.weight'], self.sd.get(k + '.bias')

    def conv(self, name, x, stride=1, relu=True):
        w, b = self._wb(name, True)
        y = F.conv2d(x, w, b, stride, (w.shape[-1] - 1) // 2)
        return F.leaky_relu(y, 0.1) if relu else y

    def deconv(self, name, x):
        w, b = self._wb(name, True)
        return F.leaky_relu(F.conv_transpose2d(x, w, b, 2, 1), 0.1)

    def flow(self, name, x):
        w, b = self._wb(name, False)
        return F.conv2d(x, w, b, 1, 1)

    def up(self, name, x):
        w, b = self._wb(name, False)
        return F.conv_transpose2d(x, w, b, 2, 1)

    def refine(self, c2, c3, c4, c5, c6, inter):
        """levels 5..2 (FlowNetC.py:104-127, FlowNetS.py:68-91, FlowNetSD.py:64-98) -> flow2"""
        flow, feat = self.flow('predict_flow6', c6), c6
        for l, enc in ((5, c5), (4, c4), (3, c3), (2, c2)):
            cat = torch.cat((enc, self.deconv('deconv%d' % l, feat), self.up('upsampled_flow%d_to_%d' % (l + 1, l), flow)), 1)
            flow = self.flow('predict_flow%d' % l, self.conv('inter_conv%d' % l, cat, relu=False) if inter else cat)
            feat = cat
        return flow


def flownet_c(sd, prefix, x):
    n = _Net(sd, prefix)
    a1, b1 = n.conv('conv1', x[:, :3], 2), n.conv('conv1', x[:, 3:], 2)
    a2, b2 = n.conv('conv2', a1, 2), n.conv('conv2', b1, 2)
    a3, b3 = n.conv('conv3', a2, 2), n.conv('conv3', b2, 2)
    corr = F.leaky_relu(correlation(a3.contiguous(), b3.contiguous()), 0.1)
    c31 = n.conv('conv3_1', torch.cat((n.conv('conv_redir', a3), corr), 1))
    c4 = n.conv('conv4_1', n.conv('conv4', c31, 2))
    c5 = n.conv('conv5_1', n.conv('conv5', c4, 2))
    c6 = n.conv('conv6_1', n.conv('conv6', c5, 2))
    return n.refine(a2, c31, c4, c5, c6, False)


def flownet_s(sd, prefix, x):
    n = _Net(sd, prefix)
    c2 = n.conv('conv2', n.conv('conv1', x, 2), 2)
    c3 = n.conv('conv3_1', n.conv('conv3', c2, 2))
    c4 = n.conv('conv4_1', n.conv('conv4', c3, 2))
    c5 = n.conv('conv5_1', n.conv('conv5', c4, 2))
    c6 = n.conv('conv6_1', n.conv('conv6', c5, 2))
    return n.refine(c2, c3, c4, c5, c6, False)


def flownet_sd(sd, prefix, x):
    n = _Net(sd, prefix)
    c1 = n.conv('conv1_1', n.conv('conv1', n.conv('conv0', x), 2))
    c2 = n.conv('conv2_1', n.conv('conv2', c1, 2))
    c3 = n.conv('conv3_1', n.conv('conv3', c2, 2))
    c4 = n.conv('conv4_1', n.conv('conv4', c3, 2))
    c5 = n.conv('conv5_1', n.conv('conv5', c4, 2))
    c6 = n.conv('conv6_1', n.conv('conv6', c5, 2))
    return n.refine(c2, c3, c4, c5, c6, True)


def flownet_fusion(sd, prefix, x):
    n = _Net(sd, prefix)
    c0 = n.conv('conv0', x)
    c1 = n.conv('conv1_1', n.conv('conv1', c0, 2))
    c2 = n.conv('conv2_1', n.conv('conv2', c1, 2))
    f2 = n.flow('predict_flow2', c2)
    cat1 = torch.cat((c1, n.deconv('deconv1', c2), n.up('upsampled_flow2_to_1', f2)), 1)
    f1 = n.flow('predict_flow1', n.conv('inter_conv1', cat1, relu=False))
    cat0 = torch.cat((c0, n.deconv('deconv0', cat1), n.up('upsampled_flow1_to_0', f1)), 1)
    return n.flow('predict_flow0', n.conv('inter_conv0', cat0, relu=False))


@torch.no_grad()
def flownet2_forward(sd, inputs, rgb_max=255., div_flow=20.):
    """inputs [B,3,2,H,W] -> (flow [B,2,H,W], intermediates)            flownet2.py:65-149"""
    sd = {k: v.float() for k, v in sd.items()}
    mean = inputs.contiguous().view(inputs.shape[:2] + (-1,)).mean(-1).view(inputs.shape[:2] + (1, 1, 1))
    x = (inputs - mean) / rgb_max
    x = torch.cat((x[:, :, 0], x[:, :, 1]), 1)
    img0, img1 = x[:, :3].contiguous(), x[:, 3:].contiguous()

    def up4(f, mode):
        return F.interpolate(f, scale_factor=4, mode=mode, **({'align_corners': False} if mode == 'bilinear' else {}))

    def stage(flow):
        warped = resample(img1, flow)
        return torch.cat((x, warped, flow / div_flow, channelnorm(img0 - warped)), 1)
    c2 = flownet_c(sd, 'flownetc.', x)
    c_flow = up4(c2 * div_flow, 'bilinear')
    s1_2 = flownet_s(sd, 'flownets_1.', stage(c_flow))
    s1_flow = up4(s1_2 * div_flow, 'bilinear')
    s2_2 = flownet_s(sd, 'flownets_2.', stage(s1_flow))
    s2_flow = up4(s2_2 * div_flow, 'nearest')
    sd_2 = flownet_sd(sd, 'flownets_d.', x)
    sd_flow = up4(sd_2 / div_flow, 'nearest')
    cat3 = torch.cat((img0, sd_flow, s2_flow, channelnorm(sd_flow), channelnorm(s2_flow),
                      channelnorm(img0 - resample(img1, sd_flow)), channelnorm(img0 - resample(img1, s2_flow))), 1)
    out = flownet_fusion(sd, 'flownetfusion.', cat3)
    return out, dict(x=x, flownetc_flow2=c2, flownets1_flow2=s1_2, flownets2_flow2=s2_2, flownetsd_flow2=sd_2, concat3=cat3)


def seeded_state(keys_shapes, seed=2026):
    """Deterministic weights for a list of (key, shape): He-style normal weights scaled so activations stay O(1) through the four
    stacked networks, small normal biases.  Replayed identically by the fixture generator (on the reference module's own
    state_dict keys) and by the tests (on ours) -- 162 M parameters do not fit a fixture, their recipe does."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in keys_shapes:
        shape = tuple(int(s) for s in shape)
        if k.endswith('.weight'):
            transposed = 'deconv' in k or 'upsampled_flow' in k
            fan_in = (shape[0] if transposed else shape[1]) * shape[2] * shape[3] / (4.0 if transposed else 1.0)
            sd[k] = torch.randn(shape, generator=g) * (1.6 / fan_in) ** 0.5
        else:
            sd[k] = torch.randn(shape, generator=g) * 0.05
    return sd
