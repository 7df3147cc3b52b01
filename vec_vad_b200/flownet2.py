"""FlowNet2 inference graph on libvecvad.so -- the network ``calc_optical_flow.py:15-22,57`` runs to
produce the optical flow the completion UNets regress.

Same surface as the reference (``FlowNet2_src/models/flownet2.py:10-149``): ``FlowNet2(with_bn=False, rgb_max=255., div_flow=20.)``,
``forward(inputs [B,3,2,H,W]) -> flow [B,2,H,W]`` (H, W multiples of 64), ``state_dict()`` keys / shapes / order equal to the
reference's, so ``FlowNet2_checkpoint.pth.tar`` loads as it does there (calc_optical_flow.py:16-21).  Sub-networks:

  flownetc      components/FlowNetC.py:10-132     two-stream encoder + correlation + refinement
  flownets_1/2  components/FlowNetS.py:9-96       12-channel stacked encoder / refinement
  flownets_d    components/FlowNetSD.py:9-103     small-displacement net (3x3 stem, inter-convs before each flow prediction)
  flownetfusion components/FlowNetFusion.py:9-64  full-resolution fusion of the two estimates

What differs from the reference is HOW it runs: the networks are tables of layers interpreted over the C ABI
(vecvad_fn_conv2d / vecvad_fn_deconv4x4s2 / ..., csrc/flownet_ops.cu, plus the three native ops), every torch.cat of the reference
is a channel slice of a pre-sized buffer that the producing kernel writes directly, the warp + difference + channel norm at each
stage boundary is one kernel, and there is no BatchNorm branch (with_bn=False is the only configuration the reference
instantiates: flownet2.py:13, calc_optical_flow.py:15).  Inference only; no CPU path.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from . import flow_ops as ops

LEAKY = 0.1                                    # misc.py:26,38


# ------------------------------------------------------------------------------------------------------------------ layer tables
# (attribute name, kind, c_in, c_out, kernel, stride).  kind: 'conv' = Conv2d + LeakyReLU wrapped in a Sequential (misc.py:6-27),
# 'lin' = the same without activation (with_relu=False), 'deconv' = ConvTranspose2d(4,2,1) + LeakyReLU in a Sequential
# (misc.py:30-38), 'flow' = bare Conv2d(c, 2, 3, 1, 1) (misc.py:41-43), 'upflow' / 'upflow_nobias' = bare ConvTranspose2d(2, 2, 4, 2, 1).
def _refine_tail(prefix_inter, bias_up):
    t = [('deconv5', 'deconv', 1024, 512, 4, 2), ('deconv4', 'deconv', 1026, 256, 4, 2), ('deconv3', 'deconv', 770, 128, 4, 2),
         ('deconv2', 'deconv', 386, 64, 4, 2)]
    if prefix_inter:
        t += [('inter_conv5', 'lin', 1026, 512, 3, 1), ('inter_conv4', 'lin', 770, 256, 3, 1), ('inter_conv3', 'lin', 386, 128, 3, 1),
              ('inter_conv2', 'lin', 194, 64, 3, 1)]
        pf = [1024, 512, 256, 128, 64]
    else:
        pf = [1024, 1026, 770, 386, 194]
    t += [('predict_flow%d' % (6 - i), 'flow', c, 2, 3, 1) for i, c in enumerate(pf)]
    up = 'upflow' if bias_up else 'upflow_nobias'
    t += [('upsampled_flow%d_to_%d' % (l, l - 1), up, 2, 2, 4, 2) for l in (6, 5, 4, 3)]
    return t


def _deep(c3_in):
    return [('conv3_1', 'conv', c3_in, 256, 3, 1), ('conv4', 'conv', 256, 512, 3, 2), ('conv4_1', 'conv', 512, 512, 3, 1),
            ('conv5', 'conv', 512, 512, 3, 2), ('conv5_1', 'conv', 512, 512, 3, 1), ('conv6', 'conv', 512, 1024, 3, 2),
            ('conv6_1', 'conv', 1024, 1024, 3, 1)]


TABLES = {
    'C': [('conv1', 'conv', 3, 64, 7, 2), ('conv2', 'conv', 64, 128, 5, 2), ('conv3', 'conv', 128, 256, 5, 2),
          ('conv_redir', 'conv', 256, 32, 1, 1)] + _deep(473) + _refine_tail(False, True),
    'S': [('conv1', 'conv', 12, 64, 7, 2), ('conv2', 'conv', 64, 128, 5, 2), ('conv3', 'conv', 128, 256, 5, 2)] + _deep(256)
         + _refine_tail(False, False),
    'SD': [('conv0', 'conv', 6, 64, 3, 1), ('conv1', 'conv', 64, 64, 3, 2), ('conv1_1', 'conv', 64, 128, 3, 1),
           ('conv2', 'conv', 128, 128, 3, 2), ('conv2_1', 'conv', 128, 128, 3, 1), ('conv3', 'conv', 128, 256, 3, 2)] + _deep(256)
          + _refine_tail(True, True),
    'Fusion': [('conv0', 'conv', 11, 64, 3, 1), ('conv1', 'conv', 64, 64, 3, 2), ('conv1_1', 'conv', 64, 128, 3, 1),
               ('conv2', 'conv', 128, 128, 3, 2), ('conv2_1', 'conv', 128, 128, 3, 1),
               ('deconv1', 'deconv', 128, 32, 4, 2), ('deconv0', 'deconv', 162, 16, 4, 2),
               ('inter_conv1', 'lin', 162, 32, 3, 1), ('inter_conv0', 'lin', 82, 16, 3, 1),
               ('predict_flow2', 'flow', 128, 2, 3, 1), ('predict_flow1', 'flow', 32, 2, 3, 1), ('predict_flow0', 'flow', 16, 2, 3, 1),
               ('upsampled_flow2_to_1', 'upflow', 2, 2, 4, 2), ('upsampled_flow1_to_0', 'upflow', 2, 2, 4, 2)],
}


class _Layer(nn.Module):
    """Parameter holder of one layer: ``weight`` (+ ``bias``) in PyTorch's own layout, initialised like flownet2.py:53-62
    (xavier_uniform weights, U(0,1) biases).  Never called: the graph interpreter reads the tensors."""

    def __init__(self, kind, cin, cout, k):
        super().__init__()
        shape = (cin, cout, k, k) if kind in ('deconv', 'upflow', 'upflow_nobias') else (cout, cin, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        nn.init.xavier_uniform_(self.weight)
        if kind != 'upflow_nobias':
            self.bias = nn.Parameter(torch.empty(cout))
            nn.init.uniform_(self.bias)
        else:
            self.register_parameter('bias', None)


class _Wrapped(nn.Module):
    """The reference wraps conv / deconv (+ activation) in nn.Sequential: the parameters live under child '0'."""

    def __init__(self, layer):
        super().__init__()
        self.add_module('0', layer)


class View:
    """Channels [c0, c1) of an NCHW buffer."""

    def __init__(self, t, c0=0, c1=None):
        self.t, self.c0, self.c1 = t, c0, t.shape[1] if c1 is None else c1

    @property
    def C(self):
        return self.c1 - self.c0

    @property
    def H(self):
        return self.t.shape[2]

    @property
    def W(self):
        return self.t.shape[3]

    @property
    def B(self):
        return self.t.shape[0]

    @property
    def ptr(self):
        return C.c_void_p(self.t.data_ptr() + 4 * self.c0 * self.H * self.W)

    @property
    def bs(self):
        return self.t.shape[1] * self.H * self.W

    def dense(self):
        """The slice as a contiguous tensor (a view when it already is one)."""
        if self.c0 == 0 and self.c1 == self.t.shape[1]:
            return self.t
        if self.B == 1:
            return self.t[:, self.c0:self.c1]
        out = self.t.new_empty((self.B, self.C, self.H, self.W))
        scale_copy(self, View(out))
        return out


def scale_copy(src, dst, mul=1.0, leaky=-1.0):
    assert src.C == dst.C and src.H == dst.H and src.W == dst.W and src.B == dst.B
    _lib.check(_lib.lib().vecvad_fn_scale_copy(src.ptr, src.bs, dst.ptr, dst.bs, src.C * src.H * src.W, float(mul), float(leaky), src.B,
                                               _lib.cur_stream()), 'fn_scale_copy')


def upsample4(src, dst, mode, mul):
    assert dst.H == 4 * src.H and dst.W == 4 * src.W and dst.C == src.C
    _lib.check(_lib.lib().vecvad_fn_upsample4(src.ptr, src.bs, src.C, src.H, src.W, dst.ptr, dst.bs, {'bilinear': 0, 'nearest': 1}[mode],
                                              float(mul), src.B, _lib.cur_stream()), 'fn_upsample4')


_SCRATCH = {}


def _scratch(device):
    """Partial sums of the layers whose contraction is split over CTAs (csrc/flownet_ops.cu plan_conv): one 64 MB buffer per
    device (VECVAD_FN_SCRATCH_MB), reused by every layer (all launches are ordered on the caller's stream).  Its size bounds the
    splits plan_conv may choose, so it is part of what fixes the summation order."""
    key = (device.type, device.index)
    if key not in _SCRATCH:
        mb = int(os.environ.get('VECVAD_FN_SCRATCH_MB', '64'))
        _SCRATCH[key] = torch.empty(mb << 18, dtype=torch.float32, device=device)
    return _SCRATCH[key]


class _SubNet(nn.Module):
    """One of FlowNetC / FlowNetS / FlowNetSD / FlowNetFusion: the parameter tree (reference attribute names and order) plus the
    interpreter helpers."""

    def __init__(self, kind):
        super().__init__()
        self.kind = kind
        self.spec = {}
        for (name, lk, cin, cout, k, s) in TABLES[kind]:
            layer = _Layer(lk, cin, cout, k)
            self.add_module(name, _Wrapped(layer) if lk in ('conv', 'lin', 'deconv') else layer)
            self.spec[name] = (lk, cin, cout, k, s, layer)
        self._phases = {}                      # deconv weights re-laid out per output parity phase (rebuilt when the weight changes)

    # -- layer execution
    def conv(self, name, src, dst=None):
        lk, cin, cout, k, s, layer = self.spec[name]
        assert lk in ('conv', 'lin', 'flow') and src.C == cin, (name, src.C, cin)
        oh, ow = (src.H + 2 * ((k - 1) // 2) - k) // s + 1, (src.W + 2 * ((k - 1) // 2) - k) // s + 1
        if dst is None:
            dst = View(src.t.new_empty((src.B, cout, oh, ow)))
        assert dst.C == cout and dst.H == oh and dst.W == ow, (name, dst.C, dst.H, dst.W)
        sc = _scratch(src.t.device)
        _lib.check(_lib.lib().vecvad_fn_conv2d(src.ptr, src.bs, cin, src.H, src.W, _lib.ptr(layer.weight), _lib.ptr(layer.bias), dst.ptr,
                                               dst.bs, cout, k, s, int(lk == 'conv'), src.B, _lib.ptr(sc), sc.numel(), _lib.cur_stream()),
                   'fn_conv2d ' + name)
        return dst

    def _phase_weights(self, name, layer):
        w = layer.weight
        key = (w.data_ptr(), w._version)
        hit = self._phases.get(name)
        if hit is None or hit[0] != key:
            taps = (C.c_int * 4)()
            _lib.check(_lib.lib().vecvad_fn_deconv_taps(taps), 'fn_deconv_taps')
            ph = []
            for py in range(2):
                for px in range(2):
                    ky = torch.tensor([taps[py * 2], taps[py * 2 + 1]], device=w.device)
                    kx = torch.tensor([taps[px * 2], taps[px * 2 + 1]], device=w.device)
                    sub = w.detach().index_select(2, ky).index_select(3, kx)           # [ci][co][2][2]
                    ph.append(sub.permute(1, 0, 2, 3).reshape(w.shape[1], -1))           # [co][ci*4]
            hit = (key, torch.stack(ph).contiguous())
            self._phases[name] = hit
        return hit[1]

    def deconv(self, name, src, dst=None):
        lk, cin, cout, k, s, layer = self.spec[name]
        assert lk in ('deconv', 'upflow', 'upflow_nobias') and src.C == cin, (name, src.C, cin)
        if dst is None:
            dst = View(src.t.new_empty((src.B, cout, 2 * src.H, 2 * src.W)))
        assert dst.C == cout and dst.H == 2 * src.H and dst.W == 2 * src.W, name
        wp = self._phase_weights(name, layer)
        sc = _scratch(src.t.device)
        _lib.check(_lib.lib().vecvad_fn_deconv4x4s2(src.ptr, src.bs, cin, src.H, src.W, _lib.ptr(wp), _lib.ptr(layer.bias), dst.ptr, dst.bs,
                                                    cout, int(lk == 'deconv'), src.B, _lib.ptr(sc), sc.numel(), _lib.cur_stream()),
                   'fn_deconv4x4s2 ' + name)
        return dst

    # -- the shared refinement: levels 5..2 of FlowNetC / S / SD (FlowNetC.py:104-127, FlowNetS.py:68-91, FlowNetSD.py:64-98)
    def refine(self, top, cats):
        """top: out_conv6; cats[l]: concat buffer of level l whose first channels already hold the encoder feature."""
        inter = self.kind == 'SD'
        flow = self.conv('predict_flow6', top)
        feat = top
        for l in (5, 4, 3, 2):
            cat = cats[l]
            n_enc = cat.shape[1] - self.spec['deconv%d' % l][2] - 2
            n_dec = self.spec['deconv%d' % l][2]
            self.deconv('upsampled_flow%d_to_%d' % (l + 1, l), flow, View(cat, n_enc + n_dec, n_enc + n_dec + 2))
            self.deconv('deconv%d' % l, feat, View(cat, n_enc, n_enc + n_dec))
            feat = View(cat)
            flow = self.conv('predict_flow%d' % l, self.conv('inter_conv%d' % l, feat) if inter else feat)
        return flow


def _buf(like, c, h, w):
    return like.new_empty((like.shape[0], c, h, w))


class FlowNet2(nn.Module):
    def __init__(self, with_bn=False, fp16=False, rgb_max=255., div_flow=20., grads=None):
        super().__init__()
        if with_bn or fp16:
            raise NotImplementedError('vec_vad_b200.FlowNet2: with_bn / fp16 are never enabled by the reference pipeline (calc_optical_flow.py:15)')
        self.with_bn, self.div_flow, self.rgb_max = with_bn, div_flow, rgb_max
        self.grads = {} if grads is None else grads
        self.flownetc = _SubNet('C')
        self.flownets_1 = _SubNet('S')
        self.flownets_2 = _SubNet('S')
        self.flownets_d = _SubNet('SD')
        self.flownetfusion = _SubNet('Fusion')
        self.corr = ops.Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)   # FlowNetC.py:24-30

    # ---- sub-networks
    def _run_c(self, x):
        n = self.flownetc
        B, _, H, W = x.shape
        t = x
        cat2, cat3 = _buf(t, 194, H // 4, W // 4), _buf(t, 386, H // 8, W // 8)
        cat4, cat5 = _buf(t, 770, H // 16, W // 16), _buf(t, 1026, H // 32, W // 32)
        c3 = []
        for img, keep in ((View(x, 0, 3), True), (View(x, 3, 6), False)):           # the two streams share conv1..conv3
            c1 = n.conv('conv1', img)
            c2 = n.conv('conv2', c1, View(cat2, 0, 128) if keep else None)
            c3.append(n.conv('conv3', c2))
        in31 = _buf(t, 473, H // 8, W // 8)
        corr = self.corr(c3[0].t, c3[1].t)                                          # [B,441,H/8,W/8]
        scale_copy(View(corr), View(in31, 32, 473), 1.0, LEAKY)                     # corr_activation (FlowNetC.py:33,91)
        n.conv('conv_redir', c3[0], View(in31, 0, 32))
        n.conv('conv3_1', View(in31), View(cat3, 0, 256))
        n.conv('conv4_1', n.conv('conv4', View(cat3, 0, 256)), View(cat4, 0, 512))
        n.conv('conv5_1', n.conv('conv5', View(cat4, 0, 512)), View(cat5, 0, 512))
        top = n.conv('conv6_1', n.conv('conv6', View(cat5, 0, 512)))
        return n.refine(top, {5: cat5, 4: cat4, 3: cat3, 2: cat2})

    def _run_s(self, n, x):
        B, _, H, W = x.shape
        cat2, cat3 = _buf(x, 194, H // 4, W // 4), _buf(x, 386, H // 8, W // 8)
        cat4, cat5 = _buf(x, 770, H // 16, W // 16), _buf(x, 1026, H // 32, W // 32)
        if n.kind == 'S':
            n.conv('conv2', n.conv('conv1', View(x)), View(cat2, 0, 128))
        else:                                                                        # FlowNetSD stem (FlowNetSD.py:55-57)
            c1 = n.conv('conv1_1', n.conv('conv1', n.conv('conv0', View(x))))
            n.conv('conv2_1', n.conv('conv2', c1), View(cat2, 0, 128))
        n.conv('conv3_1', n.conv('conv3', View(cat2, 0, 128)), View(cat3, 0, 256))
        n.conv('conv4_1', n.conv('conv4', View(cat3, 0, 256)), View(cat4, 0, 512))
        n.conv('conv5_1', n.conv('conv5', View(cat4, 0, 512)), View(cat5, 0, 512))
        top = n.conv('conv6_1', n.conv('conv6', View(cat5, 0, 512)))
        return n.refine(top, {5: cat5, 4: cat4, 3: cat3, 2: cat2})

    def _run_fusion(self, x):
        n = self.flownetfusion
        B, _, H, W = x.shape
        cat0, cat1 = _buf(x, 82, H, W), _buf(x, 162, H // 2, W // 2)
        n.conv('conv0', View(x), View(cat0, 0, 64))
        n.conv('conv1_1', n.conv('conv1', View(cat0, 0, 64)), View(cat1, 0, 128))
        c2 = n.conv('conv2_1', n.conv('conv2', View(cat1, 0, 128)))
        flow2 = n.conv('predict_flow2', c2)
        n.deconv('upsampled_flow2_to_1', flow2, View(cat1, 160, 162))
        n.deconv('deconv1', c2, View(cat1, 128, 160))
        flow1 = n.conv('predict_flow1', n.conv('inter_conv1', View(cat1)))
        n.deconv('upsampled_flow1_to_0', flow1, View(cat0, 80, 82))
        n.deconv('deconv0', View(cat1), View(cat0, 64, 80))
        return n.conv('predict_flow0', n.conv('inter_conv0', View(cat0)))

    # ---- the stack (flownet2.py:65-149)
    @torch.no_grad()
    def forward(self, inputs, return_intermediates=False):
        _lib.require_cuda(inputs)
        if inputs.dim() != 5 or inputs.shape[1] != 3 or inputs.shape[2] != 2:
            raise ValueError('FlowNet2: inputs must be [B,3,2,H,W], got %s' % (tuple(inputs.shape),))
        B, _, _, H, W = inputs.shape
        if H % 64 or W % 64:
            raise ValueError('FlowNet2: H and W must be multiples of 64 (six stride-2 stages), got %dx%d' % (H, W))
        ims = inputs.contiguous().float()
        L = _lib.lib()
        x = ims.new_empty((B, 6, H, W))
        scratch = torch.empty(3 * B, dtype=torch.float64, device=ims.device)
        _lib.check(L.vecvad_fn_normalize_pair(_lib.ptr(ims), _lib.ptr(x), _lib.ptr(scratch), B, H, W, float(self.rgb_max), _lib.cur_stream()),
                   'fn_normalize_pair')
        img0, img1 = View(x, 0, 3).dense(), View(x, 3, 6).dense()
        inter = {}

        def stage_input(flow_full):
            """cat(x, warp(img1, flow), flow / div_flow, ||img0 - warp||)  -- flownet2.py:79-87,93-101"""
            warped, _, norm = ops.warp_diff_norm(img0, img1, flow_full, want_diff=False)
            cat = ims.new_empty((B, 12, H, W))
            scale_copy(View(x), View(cat, 0, 6))
            scale_copy(View(warped), View(cat, 6, 9))
            scale_copy(View(flow_full), View(cat, 9, 11), 1.0 / self.div_flow)
            scale_copy(View(norm), View(cat, 11, 12))
            return cat

        def full_res(flow2, mode, mul):
            out = ims.new_empty((B, 2, H, W))
            upsample4(flow2, View(out), mode, mul)
            return out

        c_flow2 = self._run_c(x)
        c_flow = full_res(c_flow2, 'bilinear', self.div_flow)
        s1_flow2 = self._run_s(self.flownets_1, stage_input(c_flow))
        s1_flow = full_res(s1_flow2, 'bilinear', self.div_flow)
        s2_flow2 = self._run_s(self.flownets_2, stage_input(s1_flow))
        s2_flow = full_res(s2_flow2, 'nearest', self.div_flow)                      # upsample4 (flownet2.py:105)
        sd_flow2 = self._run_s(self.flownets_d, x)
        sd_flow = full_res(sd_flow2, 'nearest', 1.0 / self.div_flow)                # upsample3 of flow2 / div_flow (flownet2.py:122)
        cat3 = ims.new_empty((B, 11, H, W))                                         # flownet2.py:138-143
        scale_copy(View(img0), View(cat3, 0, 3))
        scale_copy(View(sd_flow), View(cat3, 3, 5))
        scale_copy(View(s2_flow), View(cat3, 5, 7))
        scale_copy(View(ops.ChannelNorm()(sd_flow)), View(cat3, 7, 8))
        scale_copy(View(ops.ChannelNorm()(s2_flow)), View(cat3, 8, 9))
        scale_copy(View(ops.warp_diff_norm(img0, img1, sd_flow, want_diff=False)[2]), View(cat3, 9, 10))
        scale_copy(View(ops.warp_diff_norm(img0, img1, s2_flow, want_diff=False)[2]), View(cat3, 10, 11))
        out = self._run_fusion(cat3).dense()
        if return_intermediates:
            inter.update(x=x, flownetc_flow2=c_flow2.dense(), flownets1_flow2=s1_flow2.dense(), flownets2_flow2=s2_flow2.dense(),
                         flownetsd_flow2=sd_flow2.dense(), concat3=cat3)
            return out, inter
        return out


def conv_flops(H, W):
    """Algorithmic FLOPs (2 x MACs) of every conv / transposed conv of one FlowNet2 forward on an HxW pair."""
    total = 0

    def walk(kind, res_of):
        t = 0
        for (name, lk, cin, cout, k, s) in TABLES[kind]:
            h, w = res_of(name)
            if lk in ('deconv', 'upflow', 'upflow_nobias'):
                t += 2 * cin * cout * 16 * h * w                # per INPUT pixel
            else:
                t += 2 * cin * cout * k * k * (h // s) * (w // s)
        return t

    def res_csd(stem):
        def f(name):
            lvl = {'conv0': 0, 'conv1': stem['conv1'], 'conv1_1': 1, 'conv2': stem['conv2'], 'conv2_1': 2, 'conv3': 2, 'conv_redir': 3, 'conv3_1': 3,
                   'conv4': 3, 'conv4_1': 4, 'conv5': 4, 'conv5_1': 5, 'conv6': 5, 'conv6_1': 6}.get(name)
            if lvl is None:
                d = [int(ch) for ch in name if ch.isdigit()]
                lvl = d[0] + 1 if name.startswith(('deconv', 'upsampled')) else d[0]
            return H >> lvl, W >> lvl
        return f
    total += walk('C', res_csd({'conv1': 0, 'conv2': 1})) + 2 * (3 * 64 * 49 * (H // 2) * (W // 2) + 64 * 128 * 25 * (H // 4) * (W // 4)
                                                                 + 128 * 256 * 25 * (H // 8) * (W // 8))      # second stream
    total += 2 * walk('S', res_csd({'conv1': 0, 'conv2': 1}))
    total += walk('SD', res_csd({'conv1': 0, 'conv2': 1}))

    def res_f(name):
        lvl = {'conv0': 0, 'conv1': 0, 'conv1_1': 1, 'conv2': 1, 'conv2_1': 2, 'deconv1': 2, 'deconv0': 1, 'inter_conv1': 1, 'inter_conv0': 0,
               'predict_flow2': 2, 'predict_flow1': 1, 'predict_flow0': 0, 'upsampled_flow2_to_1': 2, 'upsampled_flow1_to_0': 1}[name]
        return H >> lvl, W >> lvl
    total += walk('Fusion', res_f)
    return total
