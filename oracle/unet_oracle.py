"""CPU oracle for the completion-UNet hot path (TEST INFRASTRUCTURE ONLY).

This file is a plain PyTorch fp32 restatement of the reference algorithm. It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never by the
product package ``vec_vad_b200`` (which has no CPU fallback at all).

Reference behaviour restated here (paths relative to /root/reference):
  * double_conv / inconv / down / up / outconv ...... model/unet.py:4-70
  * SelfCompleteNet4 (5raw1of) ...................... model/unet.py:73-267
  * SelfCompleteNetFull (5raw5of) ................... model/unet.py:270-556
  * SelfCompleteNet1raw1of .......................... model/unet.py:559-652
  * train-step body (swapped-arg MSE, Adam eps 1e-7)  train.py:241,375-402
  * per-cube sum-squared-error scoring .............. train.py:414-427, test.py:319-335

Pinning: ``tests/golden/make_golden.py`` imports the real reference from
/root/reference and dumps inputs/outputs; ``tests/test_oracle_golden.py`` checks
this restatement against those fixtures (same state_dict keys, same seeded
initial weights, same outputs / losses / gradients / post-Adam parameters).

The networks are a *set* of independent UNets sharing one input cube:
UNet ``raw_i`` sees the cube with frame ``raw_i`` erased and must reproduce that
frame; an optional flow UNet sees the same incomplete cube and regresses the
optical flow of frame ``raw_i - raw_of_offset``.
"""
import torch
import torch.nn as nn

RAW_CH = 3   # BGR channels per frame            (model/unet.py:91)
OF_CH = 2    # optical-flow channels per frame   (model/unet.py:92)


def _double_conv(cin, cout):
    # (conv3x3 pad1 + bias -> BatchNorm2d -> ReLU) x 2      model/unet.py:9-16
    class _DC(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Sequential(
                nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

        def forward(self, x):
            return self.conv(x)
    return _DC()


class _In(nn.Module):                       # model/unet.py:22-32
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _double_conv(cin, cout)

    def forward(self, x):
        return self.conv(x)


class _Down(nn.Module):                     # model/unet.py:34-44
    def __init__(self, cin, cout):
        super().__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), _double_conv(cin, cout))

    def forward(self, x):
        return self.mpconv(x)


class _Up(nn.Module):                       # model/unet.py:46-61 (bilinear=False branch, the only one used)
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, 3, stride=2, padding=1, output_padding=1)
        self.conv = _double_conv(cin, cout)

    def forward(self, deep, skip):
        return self.conv(torch.cat([skip, self.up(deep)], dim=1))


class _Out(nn.Module):                      # model/unet.py:63-70
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.conv(x)


def _unet_forward(mods, x):
    """One UNet: inc -> down x3 -> up x3 -> outc          model/unet.py:187-196"""
    inc, d1, d2, d3, u1, u2, u3, outc = mods
    x1 = inc(x)
    x2 = d1(x1)
    x3 = d2(x2)
    x4 = d3(x3)
    y = u1(x4, x3)
    y = u2(y, x2)
    y = u3(y, x1)
    return outc(y)


class CompletionNetOracle(nn.Module):
    """Generic restatement covering all three reference classes.

    kind = 'net4'  -> SelfCompleteNet4      (attribute names inc{i}, down{i}{k}, up{i}{k}, outc{i},
                                             inc_of, down_of{k}, up_of{k}, outc_of)
    kind = 'full'  -> SelfCompleteNetFull   (flow nets named inc_of{j}, down_of{j}{k}, ...)
    kind = '1raw1of' -> SelfCompleteNet1raw1of (inc, down{k}, up{k}, outc, inc_of, ...)

    Sub-modules are registered in the same order as the reference constructors so
    that (a) ``state_dict()`` key order matches and (b) a seeded construction
    consumes the RNG identically (same initial weights).
    """

    def __init__(self, kind='net4', features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict',
                 rawRange=None, useFlow=True, padding=True):
        super().__init__()
        assert kind in ('net4', 'full', '1raw1of')
        assert tot_of_num <= tot_raw_num
        predict_modes = ('predict', 'elasticPredict') if kind == 'full' else ('predict',)
        if border_mode in predict_modes:                      # model/unet.py:78-83, 274-279
            raw_center, of_center = tot_raw_num - 1, tot_of_num - 1
        else:
            raw_center, of_center = (tot_raw_num - 1) // 2, (tot_of_num - 1) // 2
        if rawRange is None:                                   # model/unet.py:84-90
            self.rawRange = range(tot_raw_num)
        else:
            if rawRange < 0:
                rawRange += tot_raw_num
            assert rawRange < tot_raw_num
            self.rawRange = range(rawRange, rawRange + 1)
        self.kind = kind
        self.tot_raw_num, self.tot_of_num = tot_raw_num, tot_of_num
        self.raw_of_offset = raw_center - of_center
        assert self.raw_of_offset >= 0
        self.useFlow, self.padding = useFlow, padding
        cin = RAW_CH * (tot_raw_num if padding else tot_raw_num - 1)   # model/unet.py:100-103
        f = features_root

        def enc(prefix_inc, prefix_down):
            setattr(self, prefix_inc, _In(cin, f))
            for k in (1, 2, 3):
                setattr(self, '%s%d' % (prefix_down, k), _Down(f * 2 ** (k - 1), f * 2 ** k))

        def dec(prefix_up, prefix_out, cout):
            for k in (1, 2, 3):
                setattr(self, '%s%d' % (prefix_up, k), _Up(f * 2 ** (4 - k), f * 2 ** (3 - k)))
            setattr(self, prefix_out, _Out(f, cout))

        if kind == '1raw1of':                                  # model/unet.py:598-617
            enc('inc', 'down')
            dec('up', 'outc', RAW_CH)
            if useFlow:
                enc('inc_of', 'down_of')
                dec('up_of', 'outc_of', OF_CH)
            return
        n_raw = 5                                              # the reference hard-codes five raw UNets
        for i in range(n_raw):                                 # model/unet.py:110-133
            enc('inc%d' % i, 'down%d' % i)
        for i in range(n_raw):                                 # model/unet.py:135-158
            dec('up%d' % i, 'outc%d' % i, RAW_CH)
        if useFlow:
            if kind == 'net4':                                 # model/unet.py:161-170
                enc('inc_of', 'down_of')
                dec('up_of', 'outc_of', OF_CH)
            else:                                              # model/unet.py:359-408
                for j in range(5):
                    enc('inc_of%d' % j, 'down_of%d' % j)
                for j in range(5):
                    dec('up_of%d' % j, 'outc_of%d' % j, OF_CH)

    # -- helpers ---------------------------------------------------------------------------
    def _mods(self, inc, down, up, outc):
        g = lambda n: getattr(self, n)
        return (g(inc), g(down + '1'), g(down + '2'), g(down + '3'), g(up + '1'), g(up + '2'), g(up + '3'), g(outc))

    def raw_unet(self, i):
        if self.kind == '1raw1of':
            return self._mods('inc', 'down', 'up', 'outc')
        return self._mods('inc%d' % i, 'down%d' % i, 'up%d' % i, 'outc%d' % i)

    def of_unet(self, j):
        if self.kind == 'full':
            return self._mods('inc_of%d' % j, 'down_of%d' % j, 'up_of%d' % j, 'outc_of%d' % j)
        return self._mods('inc_of', 'down_of', 'up_of', 'outc_of')

    def forward(self, x, x_of):
        c = RAW_CH
        if self.kind == '1raw1of':                             # model/unet.py:619-652
            last = (self.tot_raw_num - 1) * c
            if self.padding:
                inc_x = x.clone()
                inc_x[:, last:] = 0
            else:
                inc_x = x[:, :last]
            raw_target = x[:, last:]
            raw_out = _unet_forward(self.raw_unet(0), inc_x)
            of_i = self.tot_raw_num - 1 - self.raw_of_offset
            of_out = _unet_forward(self.of_unet(0), inc_x)      # reference raises NameError if useFlow=False
            of_target = x_of[:, of_i * OF_CH:(of_i + 1) * OF_CH]
            return of_out, raw_out, of_target, raw_target
        raw_outs, raw_tgts, of_outs, of_tgts = [], [], [], []
        for raw_i in self.rawRange:                            # model/unet.py:178-259, 416-548
            if self.padding:
                inc_x = x.clone()
                inc_x[:, raw_i * c:(raw_i + 1) * c] = 0
            else:
                inc_x = torch.cat([x[:, :raw_i * c], x[:, (raw_i + 1) * c:]], dim=1)
            raw_tgts.append(x[:, raw_i * c:(raw_i + 1) * c])
            raw_outs.append(_unet_forward(self.raw_unet(raw_i), inc_x))
            of_i = raw_i - self.raw_of_offset
            if self.useFlow and 0 <= of_i < self.tot_of_num:
                of_outs.append(_unet_forward(self.of_unet(of_i), inc_x))
                of_tgts.append(x_of[:, of_i * OF_CH:(of_i + 1) * OF_CH])
        raw_outs, raw_tgts = torch.cat(raw_outs, 1), torch.cat(raw_tgts, 1)
        if len(of_outs) > 0:                                   # model/unet.py:263-265 (empty *list* otherwise)
            of_outs, of_tgts = torch.cat(of_outs, 1), torch.cat(of_tgts, 1)
        return of_outs, raw_outs, of_tgts, raw_tgts


def SelfCompleteNet4(**kw):
    kw.setdefault('tot_of_num', 1)
    return CompletionNetOracle('net4', **kw)


def SelfCompleteNetFull(**kw):
    kw.setdefault('tot_of_num', 5)
    return CompletionNetOracle('full', **kw)


def SelfCompleteNet1raw1of(**kw):
    kw.setdefault('features_root', 64)
    return CompletionNetOracle('1raw1of', **kw)


# ---- train step / scoring -------------------------------------------------------------------
def make_adam(model, lr=1e-3):
    """optim.Adam(cur_model.parameters(), eps=1e-7, weight_decay=0.0)          train.py:376"""
    return torch.optim.Adam(model.parameters(), lr=lr, eps=1e-7, weight_decay=0.0)


def train_step(model, opt, x, x_of, lambda_raw=1.0, lambda_of=1.0):
    """One iteration of the reference hot loop.                                train.py:383-402

    Returns (loss_raw, loss_of) as python floats (loss_of = 0.0 without flow).
    """
    mse = nn.MSELoss()
    of_out, raw_out, of_tgt, raw_tgt = model(x, x_of)
    loss_raw = mse(raw_tgt.detach(), raw_out)                  # arguments swapped on purpose (train.py:385)
    if model.useFlow:
        loss_of = mse(of_tgt.detach(), of_out)
        loss = lambda_raw * loss_raw + lambda_of * loss_of
    else:
        loss_of, loss = None, loss_raw
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss_raw.item()), (float(loss_of.item()) if loss_of is not None else 0.0)


@torch.no_grad()
def score_cubes(model, x, x_of):
    """Per-cube sum of squared error, raw and flow (model must be in eval()).  train.py:414-427"""
    of_out, raw_out, of_tgt, raw_tgt = model(x, x_of)
    raw = ((raw_tgt - raw_out) ** 2).sum(dim=(1, 2, 3))
    of = ((of_tgt - of_out) ** 2).sum(dim=(1, 2, 3)) if model.useFlow else None
    return raw, of


# ---- synthetic cubes ---------------------------------------------------------
def synthetic_cubes(n, t_of=1, seed=1234, grey=False):
    """uint8 raw cubes [n,5,32,32,3] and float32 flow cubes [n,t_of,32,32,2] (numpy)."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    if grey:
        raw = torch.randint(0, 256, (n, 5, 32, 32, 1), generator=g, dtype=torch.uint8).expand(-1, -1, -1, -1, 3)
    else:
        raw = torch.randint(0, 256, (n, 5, 32, 32, 3), generator=g, dtype=torch.uint8)
    flow = torch.randn((n, t_of, 32, 32, 2), generator=g, dtype=torch.float32)
    return np.ascontiguousarray(raw.numpy()), np.ascontiguousarray(flow.numpy())


def cubes_to_tensors(raw_u8, flow_f32):
    """Restatement of cube_to_train_dataset + default collate.     vad_datasets.py:130-168

    [N,T,H,W,C] -> transpose [1,2,0,3] per item -> [H,W,T*C] -> ToTensor -> [T*C,H,W];
    uint8 is scaled by 1/255 (torchvision ToTensor), float32 is passed through.
    """
    r = torch.from_numpy(raw_u8).permute(0, 1, 4, 2, 3)        # [N,T,C,H,W]
    r = r.reshape(r.shape[0], -1, r.shape[3], r.shape[4]).to(torch.float32).div(255)
    f = torch.from_numpy(flow_f32).permute(0, 1, 4, 2, 3)
    f = f.reshape(f.shape[0], -1, f.shape[3], f.shape[4]).contiguous()
    return r.contiguous(), f
