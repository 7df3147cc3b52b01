"""Drop-in module surface of the reference's ``model/unet.py`` on top of libvecvad.so.

Same class names, constructor arguments, ``forward(x, x_of)`` 4-tuple and ``state_dict`` keys as
the reference (model/unet.py:73-267 ``SelfCompleteNet4``, :270-556 ``SelfCompleteNetFull``,
:559-652 ``SelfCompleteNet1raw1of``), so ``train.py`` / ``test.py`` and checkpoints trained with
either implementation interoperate.  No arithmetic happens in PyTorch: parameters are views of
one flat device buffer that the CUDA engine reads, and ``forward`` / ``backward`` are single calls
through the C ABI (include/vecvad.h).  There is no CPU path.

Two ways to train:
  * the reference's own loop (``loss.backward(); optimizer.step()``) -- autograd sees one node;
  * ``CompletionNet.train_step`` -- forward + MSE + backward (+ gradient all-reduce) + Adam fused
    on the device, no host synchronisation (the fast path ``bench.py`` measures).
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib

RAW_CH, OF_CH = 3, 2


class _Node(nn.Module):
    """Pure naming container: reproduces the reference's module tree so state_dict keys match."""

    def child(self, name):
        if name not in self._modules:
            self.add_module(name, _Node())
        return self._modules[name]


def _unit_shapes(F, cin):
    """(C_in, C_out) of the 14 conv3x3+BN+ReLU units in execution order (model/unet.py:187-196)."""
    return [(cin, F), (F, F), (F, 2 * F), (2 * F, 2 * F), (2 * F, 4 * F), (4 * F, 4 * F), (4 * F, 8 * F), (8 * F, 8 * F),
            (8 * F, 4 * F), (4 * F, 4 * F), (4 * F, 2 * F), (2 * F, 2 * F), (2 * F, F), (F, F)]


def slot_layout(F, cin):
    """Offsets (in floats) of every tensor inside one UNet slot of the flat parameter buffer.

    Each tensor keeps PyTorch's own layout so nn.Parameter views alias the buffer directly.
    Returns (entries, param_stride, stat_entries, stat_stride); entries are
    (kind, index, field, offset, shape).
    """
    ent, off = [], 0

    def add(kind, idx, field, shape):
        nonlocal off
        n = 1
        for s in shape:
            n *= s
        ent.append((kind, idx, field, off, tuple(shape)))
        off += (n + 3) // 4 * 4          # keep every tensor 16-byte aligned
    for u, (ci, co) in enumerate(_unit_shapes(F, cin)):
        add('unit', u, 'conv_w', (co, ci, 3, 3))
        add('unit', u, 'conv_b', (co,))
        add('unit', u, 'bn_w', (co,))
        add('unit', u, 'bn_b', (co,))
    for k in range(3):
        ci = F << (3 - k)
        add('up', k, 'up_w', (ci, ci // 2, 3, 3))
        add('up', k, 'up_b', (ci // 2,))
    add('out', 0, 'out_w', (4, F, 1, 1))     # 3 (raw) or 2 (flow) rows used
    add('out', 0, 'out_b', (4,))
    pstride = (off + 63) // 64 * 64
    sent, soff = [], 0
    for u, (ci, co) in enumerate(_unit_shapes(F, cin)):
        sent.append((u, 'run_mean', soff, (co,)))
        soff += co
        sent.append((u, 'run_var', soff, (co,)))
        soff += co
    return ent, pstride, sent, (soff + 63) // 64 * 64


class _NetFn(torch.autograd.Function):
    """One autograd node for the whole UNet set; parameters enter as individual leaves."""

    @staticmethod
    def forward(ctx, net, x, x_of, *params):
        raw_out, of_out = net._run_forward(x, x_of, training=net.training, sse=None)
        ctx.net, ctx.gen = net, net._gen
        if of_out is None:
            return raw_out, raw_out.new_empty(0)
        return raw_out, of_out

    @staticmethod
    def backward(ctx, g_raw, g_of):
        net = ctx.net
        if ctx.gen != net._gen:
            raise RuntimeError('vec_vad_b200: backward() after a newer forward() -- the engine keeps the activations of the last '
                               'training forward only')
        if not net.training:
            raise RuntimeError('vec_vad_b200: backward through an eval-mode forward is not supported')
        g_raw = g_raw.contiguous().float()
        g_of = g_of.contiguous().float() if net._n_of_out > 0 else None
        net._run_backward(g_raw, g_of)
        flat = net._gflat.clone()          # detach from the engine buffer (autograd may keep / accumulate into it)
        grads = [flat[o:o + n].view(s) for (o, n, s) in net._param_views]
        return (None, None, None) + tuple(grads)


class WorkspacePool:
    """One device buffer shared by several CompletionNets that are never between a training forward and its backward at the
    same time -- e.g. the per-block models ``test.py`` keeps (one model per (scene, h_block, w_block)): without it every cached
    model would own a full activation workspace (28.5 MB per cube for 5raw5of)."""

    def __init__(self):
        self.buf, self.gen = None, 0

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None                          # release before allocating the larger one
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.gen += 1
        return self.buf


class CompletionNet(nn.Module):
    """G independent completion UNets over one cube batch, computed by the sm_100a engine."""

    def __init__(self, kind, features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True,
                 padding=True, patch_size=32, use_tensor_cores=True):
        super().__init__()
        assert kind in ('net4', 'full', '1raw1of')
        assert tot_of_num <= tot_raw_num
        self._ctor = dict(features_root=features_root, tot_raw_num=tot_raw_num, tot_of_num=tot_of_num, border_mode=border_mode,
                          rawRange=rawRange, useFlow=useFlow, padding=padding, patch_size=patch_size,
                          use_tensor_cores=use_tensor_cores)
        predict_modes = ('predict', 'elasticPredict') if kind == 'full' else ('predict',)
        if border_mode in predict_modes:                                   # model/unet.py:78-83
            self.raw_center_idx, self.of_center_idx = tot_raw_num - 1, tot_of_num - 1
        else:
            self.raw_center_idx, self.of_center_idx = (tot_raw_num - 1) // 2, (tot_of_num - 1) // 2
        if rawRange is None:                                               # model/unet.py:84-90
            self.rawRange = range(tot_raw_num)
        else:
            if rawRange < 0:
                rawRange += tot_raw_num
            assert rawRange < tot_raw_num
            self.rawRange = range(rawRange, rawRange + 1)
        self.kind = kind
        self.raw_channel_num, self.of_channel_num = RAW_CH, OF_CH
        self.tot_raw_num, self.tot_of_num = tot_raw_num, tot_of_num
        self.raw_of_offset = self.raw_center_idx - self.of_center_idx
        assert self.raw_of_offset >= 0
        self.useFlow, self.padding = useFlow, padding
        self.features_root, self.patch_size = features_root, patch_size
        # contraction path: False / 0 = fp32 SIMT tiles; True / 1 / 'tf32' = tcgen05 kind::tf32 tiles; 2 / 'f16' = tcgen05 kind::f16 tiles
        # over fp16 activations / gradients / weights in HBM (fp32 accumulation, statistics, losses, optimiser)
        self.use_tensor_cores = {'tf32': 1, 'f16': 2, 'fp16': 2, 'simt': 0, 'fp32': 0}.get(use_tensor_cores, use_tensor_cores)
        self.use_tensor_cores = int(self.use_tensor_cores)
        if self.use_tensor_cores not in (0, 1, 2):
            raise ValueError('use_tensor_cores must be False, True / "tf32" or 2 / "f16"')
        if kind != '1raw1of' and tot_raw_num != 5:
            raise NotImplementedError('the reference builds exactly five raw UNets (model/unet.py:110-158)')
        cin = RAW_CH * (tot_raw_num if padding else tot_raw_num - 1)       # model/unet.py:100-103
        self._cin = cin
        F = features_root

        # ---- slots (one per UNet that owns parameters) and their reference attribute names
        slots = []   # (inc, down, up, outc, out_channels)
        if kind == '1raw1of':
            slots.append(('inc', 'down', 'up', 'outc', RAW_CH))
            if useFlow:
                slots.append(('inc_of', 'down_of', 'up_of', 'outc_of', OF_CH))
            self._raw_slot = {tot_raw_num - 1: 0}
            self._of_slot = {0: 1}
        else:
            for i in range(5):
                slots.append(('inc%d' % i, 'down%d' % i, 'up%d' % i, 'outc%d' % i, RAW_CH))
            self._raw_slot = {i: i for i in range(5)}
            self._of_slot = {}
            if useFlow:
                if kind == 'net4':
                    slots.append(('inc_of', 'down_of', 'up_of', 'outc_of', OF_CH))
                    self._of_slot = {j: 5 for j in range(tot_of_num)}     # the single flow UNet serves every of_i (model/unet.py:247-259)
                else:
                    for j in range(5):
                        slots.append(('inc_of%d' % j, 'down_of%d' % j, 'up_of%d' % j, 'outc_of%d' % j, OF_CH))
                    self._of_slot = {j: 5 + j for j in range(5)}
        self._slots = slots
        ent, self._pstride, sent, self._sstride = slot_layout(F, cin)
        self._entries, self._stat_entries = ent, sent
        nslots = len(slots)

        # ---- flat buffers (CPU at construction like any nn.Module; moved by .cuda()/.to())
        self._pflat = torch.zeros(nslots * self._pstride)
        self._gflat = None
        self._phase_views = None
        self._phase_ranges = None
        self._tail_adam = os.environ.get('VECVAD_TAIL_ADAM', '1') != '0'
        self._sflat = torch.zeros(nslots * self._sstride)
        self._nbt = torch.zeros(nslots * _lib.N_UNITS, dtype=torch.long)
        self._nbt_pending = 0
        self._init_reference_order()          # seeded-init parity with the reference constructors
        self._register_tree()
        # ---- engine state
        self._parts = None                    # [dict(net, ws, g0, g1, stream)]: the UNets split over concurrent streams
        self._ws_batch = 0
        self._ws_pool, self._ws_pool_gen = None, -1     # optional WorkspacePool shared with other nets (scoring)
        self._gen = 0
        self._adam = None
        self._plan()

    # ------------------------------------------------------------------ construction helpers
    def _slot_names(self, s):
        """[(entry index or stat tuple, dotted state_dict key)] of slot s in the reference's registration order."""
        inc, down, up, outc, _ = self._slots[s]

        def unit_prefix(u):
            if u < 2:
                return '%s.conv.conv' % inc, u
            if u < 8:
                return '%s%d.mpconv.1.conv' % (down, (u - 2) // 2 + 1), u % 2
            return '%s%d.conv.conv' % (up, (u - 8) // 2 + 1), u % 2
        names = {}
        for (kind, idx, field, off, shape) in self._entries:
            if kind == 'unit':
                pre, second = unit_prefix(idx)
                pos = {'conv_w': (0, 'weight'), 'conv_b': (0, 'bias'), 'bn_w': (1, 'weight'), 'bn_b': (1, 'bias')}[field]
                names[(kind, idx, field)] = '%s.%d.%s' % (pre, pos[0] + 3 * second, pos[1])
            elif kind == 'up':
                names[(kind, idx, field)] = '%s%d.up.%s' % (up, idx + 1, 'weight' if field == 'up_w' else 'bias')
            else:
                names[(kind, idx, field)] = '%s.conv.%s' % (outc, 'weight' if field == 'out_w' else 'bias')
        for (u, field, off, shape) in self._stat_entries:
            pre, second = unit_prefix(u)
            names[('stat', u, field)] = '%s.%d.%s' % (pre, 1 + 3 * second, 'running_mean' if field == 'run_mean' else 'running_var')
            names[('nbt', u, '')] = '%s.%d.num_batches_tracked' % (pre, 1 + 3 * second)
        return names

    def _slot_view(self, s, kind, idx, field):
        for (k, i, f, off, shape) in self._entries:
            if (k, i, f) == (kind, idx, field):
                if kind == 'out':                       # only the rows this UNet owns are parameters
                    oc = self._slots[s][4]
                    shape = (oc,) + shape[1:]
                n = 1
                for d in shape:
                    n *= d
                return s * self._pstride + off, n, shape
        raise KeyError((kind, idx, field))

    @torch.no_grad()
    def _init_reference_order(self):
        """Initialise with torch's own layer constructors, created in the reference's order, so that
        ``torch.manual_seed(k); Net(...)`` yields the same weights as the reference under the same seed."""
        F, cin = self.features_root, self._cin
        shapes = _unit_shapes(F, cin)

        def fill(s, kind, idx, field, t):
            off, n, shape = self._slot_view(s, kind, idx, field)
            self._pflat[off:off + n].copy_(t.reshape(-1))

        def init_unit(s, u):
            ci, co = shapes[u]
            conv = nn.Conv2d(ci, co, 3, padding=1)
            fill(s, 'unit', u, 'conv_w', conv.weight)
            fill(s, 'unit', u, 'conv_b', conv.bias)
            fill(s, 'unit', u, 'bn_w', torch.ones(co))
            fill(s, 'unit', u, 'bn_b', torch.zeros(co))

        def enc(s):
            for u in range(8):
                init_unit(s, u)

        def dec(s):
            for k in range(3):
                ci = F << (3 - k)
                up = nn.ConvTranspose2d(ci, ci // 2, 3, stride=2, padding=1, output_padding=1)
                fill(s, 'up', k, 'up_w', up.weight)
                fill(s, 'up', k, 'up_b', up.bias)
                init_unit(s, 8 + 2 * k)
                init_unit(s, 9 + 2 * k)
            oc = self._slots[s][4]
            out = nn.Conv2d(F, oc, 1)
            fill(s, 'out', 0, 'out_w', out.weight)
            fill(s, 'out', 0, 'out_b', out.bias)
        nraw = 1 if self.kind == '1raw1of' else 5
        if self.kind == '1raw1of':
            enc(0), dec(0)
            if self.useFlow:
                enc(1), dec(1)
        else:
            for s in range(nraw):
                enc(s)
            for s in range(nraw):
                dec(s)
            for s in range(nraw, len(self._slots)):
                enc(s)
            for s in range(nraw, len(self._slots)):
                dec(s)
        for s in range(len(self._slots)):               # running_var starts at 1
            for (u, field, off, shape) in self._stat_entries:
                if field == 'run_var':
                    self._sflat[s * self._sstride + off:s * self._sstride + off + shape[0]] = 1.0

    def _register_tree(self):
        """Create the naming tree + nn.Parameter / buffer views in the reference's registration order."""
        self._param_views = []       # (offset, numel, shape) in parameters() order
        self._param_list = []
        order = []                   # slots interleaved exactly like the reference constructors
        nraw = 1 if self.kind == '1raw1of' else 5
        if self.kind == '1raw1of':
            order = [(s, part) for s in range(len(self._slots)) for part in ('enc', 'dec')]
        else:
            order = [(s, 'enc') for s in range(nraw)] + [(s, 'dec') for s in range(nraw)]
            order += [(s, 'enc') for s in range(nraw, len(self._slots))] + [(s, 'dec') for s in range(nraw, len(self._slots))]
        self._bindings = []          # (owner module, attr, kind, flat offset, numel, shape)

        def bind(key, kind, off, n, shape):
            parts = key.split('.')
            node = self
            for i, pth in enumerate(parts[:-1]):
                if node is self:
                    if pth not in self._modules:
                        self.add_module(pth, _Node())
                    node = self._modules[pth]
                else:
                    node = node.child(pth)
            attr = parts[-1]
            if kind == 'param':
                p = nn.Parameter(self._pflat[off:off + n].view(shape))
                node.register_parameter(attr, p)
                self._param_views.append((off, n, shape))
                self._param_list.append(p)
            elif kind == 'stat':
                node.register_buffer(attr, self._sflat[off:off + n].view(shape))
            else:
                node.register_buffer(attr, self._nbt[off])
            self._bindings.append((node, attr, kind, off, n, shape))
        for (s, part) in order:
            names = self._slot_names(s)
            units = range(8) if part == 'enc' else None
            seq = []
            if part == 'enc':
                seq = [('unit', u) for u in units]
            else:
                for k in range(3):
                    seq += [('up', k), ('unit', 8 + 2 * k), ('unit', 9 + 2 * k)]
                seq += [('out', 0)]
            for (kind, idx) in seq:
                if kind == 'unit':
                    for field in ('conv_w', 'conv_b'):
                        off, n, shape = self._slot_view(s, 'unit', idx, field)
                        bind(names[('unit', idx, field)], 'param', off, n, shape)
                    for field in ('bn_w', 'bn_b'):
                        off, n, shape = self._slot_view(s, 'unit', idx, field)
                        bind(names[('unit', idx, field)], 'param', off, n, shape)
                    for (u, field, soff, shape) in self._stat_entries:
                        if u == idx:
                            bind(names[('stat', u, field)], 'stat', s * self._sstride + soff, shape[0], shape)
                    bind(names[('nbt', idx, '')], 'nbt', s * _lib.N_UNITS + idx, 1, ())
                elif kind == 'up':
                    for field in ('up_w', 'up_b'):
                        off, n, shape = self._slot_view(s, 'up', idx, field)
                        bind(names[('up', idx, field)], 'param', off, n, shape)
                else:
                    for field in ('out_w', 'out_b'):
                        off, n, shape = self._slot_view(s, 'out', 0, field)
                        bind(names[('out', 0, field)], 'param', off, n, shape)

    def _rebind(self):
        """Point every Parameter / buffer back at the flat buffers (after a device / dtype move)."""
        for (node, attr, kind, off, n, shape) in self._bindings:
            if kind == 'param':
                node._parameters[attr].data = self._pflat[off:off + n].view(shape)
                node._parameters[attr].grad = None
            elif kind == 'stat':
                node._buffers[attr] = self._sflat[off:off + n].view(shape)
            else:
                node._buffers[attr] = self._nbt[off]

    def _flush_nbt(self):
        if self._nbt_pending:
            self._nbt += self._nbt_pending
            self._nbt_pending = 0

    def state_dict(self, *args, **kwargs):
        self._flush_nbt()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        self._nbt_pending = 0                               # the loaded counters replace whatever was pending
        return super().load_state_dict(state_dict, *args, **kwargs)

    def _apply(self, fn, recurse=True):
        # Move the flat buffers, then re-create the views: the default per-tensor _apply would break the aliasing.
        new_p = fn(self._pflat)
        if new_p.dtype != torch.float32:
            raise RuntimeError('vec_vad_b200: parameters are float32 (the engine computes conv tiles in TF32/FP32)')
        self._pflat = new_p.contiguous()
        self._sflat = fn(self._sflat).contiguous()
        self._flush_nbt()
        nbt = fn(self._nbt)
        self._nbt = nbt.long() if nbt.dtype != torch.long else nbt
        self._gflat = None
        self._phase_views = None
        self._phase_ranges = None
        self._release_engine()
        self._rebind()
        return self

    # ------------------------------------------------------------------ engine plumbing
    def _plan(self):
        """Which UNets run in one forward (model/unet.py:178-259) and where their outputs go."""
        plan = []   # (slot, erase_frame, out_channels, target_is_flow, target_index, out_slot)
        n_raw = n_of = 0
        if self.kind == '1raw1of':
            last = self.tot_raw_num - 1
            plan.append((0, last, RAW_CH, 0, last, 0))
            n_raw = 1
            if self.useFlow:
                plan.append((1, last, OF_CH, 1, last - self.raw_of_offset, 0))
                n_of = 1
        else:
            for raw_i in self.rawRange:
                plan.append((self._raw_slot[raw_i], raw_i, RAW_CH, 0, raw_i, n_raw))
                n_raw += 1
                of_i = raw_i - self.raw_of_offset
                if self.useFlow and 0 <= of_i < self.tot_of_num:
                    plan.append((self._of_slot[of_i], raw_i, OF_CH, 1, of_i, n_of))
                    n_of += 1
        # a slot may appear once per forward only (its activations live in per-UNet workspace): true for every
        # reference configuration (Net4 has tot_of_num == 1).
        assert len({p[0] for p in plan}) == len(plan), 'a UNet would run twice in one forward'
        self._plan_list, self._n_raw_out, self._n_of_out = plan, n_raw, n_of

    def _config(self, g0=0, g1=None):
        plan = self._plan_list[g0:g1]
        cfg = _lib.NetConfig()
        cfg.n_unets = len(plan)
        cfg.n_raw_total, cfg.n_of_total = self._n_raw_out, self._n_of_out
        cfg.features_root, cfg.tot_raw_num, cfg.patch = self.features_root, self.tot_raw_num, self.patch_size
        cfg.padding = int(bool(self.padding))
        for g, (slot, erase, oc, isflow, tidx, oslot) in enumerate(plan):
            cfg.param_slot[g], cfg.erase_frame[g], cfg.out_channels[g] = slot, erase, oc
            cfg.target_is_flow[g], cfg.target_index[g], cfg.out_slot[g] = isflow, tidx, oslot
        cfg.slot_param_stride, cfg.slot_stat_stride = self._pstride, self._sstride
        for (kind, idx, field, off, shape) in self._entries:
            if kind == 'unit':
                getattr(cfg, field)[idx] = off
            elif kind == 'up':
                getattr(cfg, field)[idx] = off
            else:
                setattr(cfg, field, off)
        for (u, field, off, shape) in self._stat_entries:
            getattr(cfg, field)[u] = off
        cfg.use_tensor_cores = int(self.use_tensor_cores)
        return cfg

    def _release_engine(self):
        for part in (getattr(self, '_parts', None) or []):
            _lib.lib().vecvad_net_destroy(part['net'])
        self._parts, self._ws_batch = None, 0

    def __del__(self):
        try:
            self._release_engine()
        except Exception:
            pass

    def __deepcopy__(self, memo):
        # the engine handle and the flat-buffer aliasing cannot be copied field by field: rebuild, then copy the state
        with torch.random.fork_rng(devices=[]):
            new = CompletionNet(self.kind, **self._ctor)
        new.__class__ = self.__class__
        dev = self._pflat.device
        new._apply(lambda t: t.to(dev))
        with torch.no_grad():
            new._pflat.copy_(self._pflat)
            new._sflat.copy_(self._sflat)
            self._flush_nbt()
            new._nbt.copy_(self._nbt)
        new.train(self.training)
        return new

    def _split(self):
        """The G UNets are independent and MAY be split into parts that run on concurrent streams (VECVAD_NET_PARTS=k).
        Measured on B200 at batch 128 (5raw1of): 1 part 4.45 ms/step, 2 parts 5.27, 3 parts 4.73 -- the persistent
        tensor-core tiles of two parts end up co-resident and contend for the same SM ingress, so the default is one part;
        the overlap that pays is inside the engine (weight-gradient tiles on a side stream, csrc/net.cu)."""
        import os
        G = len(self._plan_list)
        want = int(os.environ.get('VECVAD_NET_PARTS', '1'))
        k = max(1, min(want, G // 2 if G >= 4 else 1))
        bounds = [round(i * G / k) for i in range(k + 1)]
        return [(bounds[i], bounds[i + 1]) for i in range(k)]

    def _engine(self, batch):
        _lib.require_cuda(self._pflat)
        L = _lib.lib()
        if self._gflat is None:
            self._gflat = torch.zeros_like(self._pflat)
        if self._parts is None:
            self._parts = []
            for i, (g0, g1) in enumerate(self._split()):
                h = C.c_void_p()
                cfg = self._config(g0, g1)
                _lib.check(L.vecvad_net_create(C.byref(cfg), C.byref(h)), 'net_create')
                self._parts.append(dict(net=h, ws=None, g0=g0, g1=g1,
                                        stream=None if i == 0 else torch.cuda.Stream(device=self._pflat.device)))
        pool = self._ws_pool if len(self._parts) == 1 else None
        if batch > self._ws_batch or (pool is not None and pool.gen != self._ws_pool_gen):
            batch = max(batch, self._ws_batch)
            for part in self._parts:
                nb = C.c_int64()
                _lib.check(L.vecvad_net_workspace_bytes(part['net'], batch, C.byref(nb)), 'workspace_bytes')
                part['ws'] = None                       # drop the old workspace BEFORE allocating its replacement
                part['ws'] = (pool.get(nb.value + 256, self._pflat.device) if pool is not None else
                              torch.empty(nb.value + 256, dtype=torch.uint8, device=self._pflat.device))
                base = (part['ws'].data_ptr() + 255) // 256 * 256
                _lib.check(L.vecvad_net_bind(part['net'], _lib.ptr(self._pflat), _lib.ptr(self._gflat), _lib.ptr(self._sflat),
                                             C.c_void_p(base), nb.value, batch), 'net_bind')
            self._ws_batch = batch
            if pool is not None:
                self._ws_pool_gen = pool.gen
        return self._parts

    def share_workspace(self, pool):
        """Use ``pool`` (a WorkspacePool) instead of a private workspace.  Only for nets that are scored, or trained one at a
        time: the pool's bytes are overwritten by whichever net runs next."""
        self._ws_pool, self._ws_pool_gen = pool, -1
        for part in (self._parts or []):
            part['ws'] = None
        self._ws_batch = 0
        return self

    def _on_parts(self, fn):
        """Run fn(part, stream_handle) for every part: part 0 on the caller's stream, the others on their own stream, forked
        after everything already queued and joined back before anything queued afterwards."""
        cur = torch.cuda.current_stream()
        start = None
        for part in self._parts:
            st = part['stream']
            if st is None:
                fn(part, C.c_void_p(cur.cuda_stream))
            else:
                if start is None:
                    start = cur.record_event()
                st.wait_event(start)
                fn(part, C.c_void_p(st.cuda_stream))
        for part in self._parts:
            if part['stream'] is not None:
                cur.wait_stream(part['stream'])

    def _run_forward(self, x, x_of, training, sse, lambda_raw=1.0, lambda_of=1.0, want_outputs=True):
        _lib.require_cuda(x, x_of if torch.is_tensor(x_of) else None)
        if x.dim() != 4 or x.shape[1] != RAW_CH * self.tot_raw_num or x.shape[2] != self.patch_size or x.shape[3] != self.patch_size:
            raise ValueError('x must be [B,%d,%d,%d], got %s' % (RAW_CH * self.tot_raw_num, self.patch_size, self.patch_size, tuple(x.shape)))
        x = x.contiguous().float()
        B = x.shape[0]
        xo, xoc = None, 0
        if self._n_of_out > 0 and torch.is_tensor(x_of):
            xo = x_of.contiguous().float()
            xoc = xo.shape[1]
        self._engine(B)
        S = self.patch_size
        raw_out = x.new_empty((B, RAW_CH * self._n_raw_out, S, S)) if want_outputs else None
        of_out = x.new_empty((B, OF_CH * self._n_of_out, S, S)) if (want_outputs and self._n_of_out > 0) else None
        self._gen += 1
        L = _lib.lib()

        def fwd(part, stream):
            sse_p = None if sse is None else C.c_void_p(sse.data_ptr() + 4 * part['g0'] * B)      # sse is [G][B]: rows g0..g1
            _lib.check(L.vecvad_net_forward(part['net'], _lib.ptr(x), _lib.ptr(xo), xoc, B, int(training), _lib.ptr(raw_out),
                                            RAW_CH * self._n_raw_out, _lib.ptr(of_out), OF_CH * self._n_of_out, sse_p,
                                            float(lambda_raw), float(lambda_of), stream), 'net_forward')
        self._on_parts(fwd)
        if training:
            self._nbt_pending += 1                          # BatchNorm2d.num_batches_tracked: counted on the host, written to the
                                                            # buffers when somebody looks (state_dict / copy / move): no kernel per step
        self._keep = (x, xo)                                # inputs must outlive the asynchronous kernels
        return raw_out, of_out

    def _run_backward(self, g_raw, g_of):
        L = _lib.lib()
        self._on_parts(lambda part, stream: _lib.check(L.vecvad_net_backward(part['net'], _lib.ptr(g_raw), _lib.ptr(g_of), stream),
                                                       'net_backward'))

    # ------------------------------------------------------------------ reference surface
    def _targets(self, x, x_of):
        c = RAW_CH
        if self.kind == '1raw1of':
            last = (self.tot_raw_num - 1) * c
            of_i = self.tot_raw_num - 1 - self.raw_of_offset
            return x[:, last:], (x_of[:, of_i * OF_CH:(of_i + 1) * OF_CH] if self.useFlow else None)
        raw = [x[:, i * c:(i + 1) * c] for i in self.rawRange]
        raw_t = raw[0] if len(raw) == 1 else (x if len(raw) == self.tot_raw_num else torch.cat(raw, 1))
        of = [x_of[:, p[4] * OF_CH:(p[4] + 1) * OF_CH] for p in self._plan_list if p[3]]
        of_t = None if not of else (of[0] if len(of) == 1 else torch.cat(of, 1))
        return raw_t, of_t

    def forward(self, x, x_of):
        """-> (of_outputs, raw_outputs, of_targets, raw_targets)            model/unet.py:267,556,652"""
        if self.kind == '1raw1of' and not self.useFlow:
            raise NotImplementedError('SelfCompleteNet1raw1of(useFlow=False).forward is undefined in the reference (model/unet.py:652)')
        need_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self._param_list)
        if need_grad:
            raw_out, of_out = _NetFn.apply(self, x, x_of, *self._param_list)
            if self._n_of_out == 0:
                of_out = None
        else:
            raw_out, of_out = self._run_forward(x, x_of, training=self.training, sse=None)
        raw_t, of_t = self._targets(x.float(), x_of.float() if torch.is_tensor(x_of) else x_of)
        if of_out is None:
            return [], raw_out, [], raw_t                   # the reference returns empty lists without flow (model/unet.py:263-267)
        return of_out, raw_out, of_t, raw_t

    @torch.no_grad()
    def score(self, x, x_of):
        """Per-cube sum of squared error (raw, flow) -- train.py:414-427 / test.py:319-335 -- reduced on the device."""
        G = len(self._plan_list)
        sse = torch.empty((G, x.shape[0]), dtype=torch.float32, device=x.device)
        self._run_forward(x, x_of, training=False, sse=sse, want_outputs=False)
        is_flow = torch.tensor([p[3] for p in self._plan_list], dtype=torch.bool, device=x.device)
        raw = sse[~is_flow].sum(0)
        of = sse[is_flow].sum(0) if self._n_of_out > 0 else None
        return raw, of

    # ------------------------------------------------------------------ fused train step
    def init_adam(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.0):
        """State of ``optim.Adam(model.parameters(), eps=1e-7, weight_decay=0.0)`` (train.py:376) on the flat buffer."""
        _lib.require_cuda(self._pflat)
        self._adam = dict(lr=lr, b1=betas[0], b2=betas[1], eps=eps, wd=weight_decay, step=0,
                          m=torch.zeros_like(self._pflat), v=torch.zeros_like(self._pflat))
        return self._adam

    def train_step(self, x, x_of, lambda_raw=1.0, lambda_of=1.0, losses=None, sse=None, reduce_grads=None):
        """forward + MSE losses + backward (+ ``reduce_grads(flat_grads)``) + Adam, all asynchronous on the current stream.

        Equivalent to train.py:383-402.  Returns a device tensor ``[loss_raw, loss_of]`` (no host sync)."""
        if self._adam is None:
            self.init_adam()
        if not self.training:
            raise RuntimeError('train_step needs model.train()')
        B, G = x.shape[0], len(self._plan_list)
        if sse is None:
            sse = torch.empty((G, B), dtype=torch.float32, device=x.device)
        if losses is None:
            losses = torch.empty(2, dtype=torch.float32, device=x.device)
        L = _lib.lib()
        self._run_forward(x, x_of, training=True, sse=sse, lambda_raw=lambda_raw, lambda_of=lambda_of, want_outputs=False)
        if len(self._parts) == 1:
            _lib.check(L.vecvad_net_losses(self._parts[0]['net'], _lib.ptr(sse), B, _lib.ptr(losses), _lib.cur_stream()), 'net_losses')
        else:                                               # each part: its share of the two means (global normalisation)
            part_losses = torch.empty((len(self._parts), 2), dtype=torch.float32, device=x.device)
            for i, part in enumerate(self._parts):
                _lib.check(L.vecvad_net_losses(part['net'], C.c_void_p(sse.data_ptr() + 4 * part['g0'] * B), B,
                                               C.c_void_p(part_losses.data_ptr() + 8 * i), _lib.cur_stream()), 'net_losses')
            torch.sum(part_losses, dim=0, out=losses)
        # the backward leaves the side stream unjoined: phases 0 / 1 of the update run behind it while the last weight gradients and their
        # scatter (phase 2, ~50 us of side-stream work after the last main-stream kernel) finish
        tail = len(self._parts) == 1 and self._tail_adam
        net0 = self._parts[0]['net']
        if tail:
            _lib.check(L.vecvad_net_defer_join(net0, 1), 'defer_join')
        try:
            self._run_backward(None, None)
        finally:
            if tail:
                _lib.check(L.vecvad_net_defer_join(net0, 0), 'defer_join')
        scale = 1.0
        if tail and reduce_grads is not None:
            _lib.check(L.vecvad_net_grad_phase_wait(net0, 2, _lib.cur_stream()), 'grad_phase_wait')      # the exchange needs every gradient
            tail = False
        if reduce_grads is not None and getattr(reduce_grads, 'shard_optimizer', False):
            # reduce-scatter of the gradients, Adam on this rank's 1/world of the parameters, all-gather of the parameters
            self._adam['step'] += 1
            reduce_grads.sharded_step(self)
            return losses
        if reduce_grads is not None:
            # a reducer with ``reduce_phased`` exchanges each gradient phase as soon as the backward has produced it (the rest of
            # the backward is still running); anything else gets the whole flat buffer once the backward is complete
            if hasattr(reduce_grads, 'reduce_phased') and len(self._parts) == 1:
                scale = float(reduce_grads.reduce_phased(self))
            else:
                scale = float(reduce_grads(self._gflat))
        a = self._adam
        a['step'] += 1
        if tail:
            if self._phase_ranges is None:
                b, e = (C.c_int64 * 3)(), (C.c_int64 * 3)()
                _lib.check(L.vecvad_net_grad_phase_ranges(net0, b, e), 'grad_phase_ranges')
                self._phase_ranges = [(int(b[i]), int(e[i])) for i in range(3)]
            for ph, (b0, e0) in enumerate(self._phase_ranges):                 # same stream: phase by phase, the join before the last
                _lib.check(L.vecvad_net_grad_phase_wait(net0, ph, _lib.cur_stream()), 'grad_phase_wait')
                _lib.check(L.vecvad_adam_step_ranges(_lib.ptr(self._pflat), _lib.ptr(self._gflat), _lib.ptr(a['m']), _lib.ptr(a['v']),
                                                     len(self._slots), self._pstride, b0, e0, a['lr'], a['b1'], a['b2'], a['eps'], a['wd'],
                                                     a['step'], scale, _lib.cur_stream()), 'adam_step_ranges')
        else:
            _lib.check(L.vecvad_adam_step(_lib.ptr(self._pflat), _lib.ptr(self._gflat), _lib.ptr(a['m']), _lib.ptr(a['v']),
                                          self._pflat.numel(), a['lr'], a['b1'], a['b2'], a['eps'], a['wd'], a['step'], scale,
                                          _lib.cur_stream()), 'adam_step')
        return losses

    def _adam_flat(self, off, n, scale):
        """torch.optim.Adam's update (train.py:376) on floats [off, off + n) of the flat buffers, current stream."""
        a = self._adam
        _lib.check(_lib.lib().vecvad_adam_step(C.c_void_p(self._pflat.data_ptr() + 4 * off), C.c_void_p(self._gflat.data_ptr() + 4 * off),
                                               C.c_void_p(a['m'].data_ptr() + 4 * off), C.c_void_p(a['v'].data_ptr() + 4 * off), n, a['lr'],
                                               a['b1'], a['b2'], a['eps'], a['wd'], a['step'], scale, _lib.cur_stream()), 'adam_step')

    def train_step_empty(self, reduce_grads):
        """A data-parallel step for a rank whose share of a ragged last batch is EMPTY: no forward / backward, a zero gradient
        into the collective, then the same Adam update as every other rank (the replicas must stay identical).  BatchNorm
        running statistics are not touched (no batch was seen), like a DataParallel replica that received no chunk."""
        if self._adam is None:
            self.init_adam()
        _lib.require_cuda(self._pflat)
        if self._gflat is None:
            self._gflat = torch.zeros_like(self._pflat)
        self._gflat.zero_()
        if reduce_grads is not None and getattr(reduce_grads, 'shard_optimizer', False):
            self._adam['step'] += 1
            reduce_grads.sharded_step(self)
            return torch.zeros(2, dtype=torch.float32, device=self._pflat.device)
        scale = float(reduce_grads(self._gflat)) if reduce_grads is not None else 1.0
        a = self._adam
        a['step'] += 1
        _lib.check(_lib.lib().vecvad_adam_step(_lib.ptr(self._pflat), _lib.ptr(self._gflat), _lib.ptr(a['m']), _lib.ptr(a['v']),
                                               self._pflat.numel(), a['lr'], a['b1'], a['b2'], a['eps'], a['wd'], a['step'], scale,
                                               _lib.cur_stream()), 'adam_step')
        return torch.zeros(2, dtype=torch.float32, device=self._pflat.device)

    # ------------------------------------------------------------------ gradient phases (overlapped data-parallel exchange)
    def grad_phase_views(self):
        """[[contiguous 1-D views of the flat gradient buffer] per phase], in the order the backward completes them: phase 0 the
        decoder, phase 1 the deepest encoder block, phase 2 the rest -- one view per UNet slot (include/vecvad.h
        vecvad_net_grad_phase_ranges).  Together the views cover every parameter gradient exactly once."""
        if self._phase_views is None:
            self._engine(max(1, self._ws_batch))
            b, e = (C.c_int64 * 3)(), (C.c_int64 * 3)()
            _lib.check(_lib.lib().vecvad_net_grad_phase_ranges(self._parts[0]['net'], b, e), 'grad_phase_ranges')
            self._phase_views = [[self._gflat[s * self._pstride + b[ph]:s * self._pstride + e[ph]] for s in range(len(self._slots))]
                                 for ph in range(3)]
        return self._phase_views

    def wait_grad_phase(self, phase, stream=None):
        """Make ``stream`` (default: the current stream) wait until phase ``phase`` of the backward queued last is complete."""
        st = torch.cuda.current_stream() if stream is None else stream
        _lib.check(_lib.lib().vecvad_net_grad_phase_wait(self._parts[0]['net'], int(phase), C.c_void_p(st.cuda_stream)), 'grad_phase_wait')

    @property
    def flat_params(self):
        return self._pflat

    @property
    def flat_grads(self):
        return self._gflat


class SelfCompleteNet4(CompletionNet):          # 5raw1of                                   model/unet.py:73
    def __init__(self, features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True,
                 padding=True, **kw):
        super().__init__('net4', features_root, tot_raw_num, tot_of_num, border_mode, rawRange, useFlow, padding, **kw)


class SelfCompleteNetFull(CompletionNet):       # 5raw5of                                   model/unet.py:270
    def __init__(self, features_root=32, tot_raw_num=5, tot_of_num=5, border_mode='predict', rawRange=None, useFlow=True,
                 padding=True, **kw):
        super().__init__('full', features_root, tot_raw_num, tot_of_num, border_mode, rawRange, useFlow, padding, **kw)


class SelfCompleteNet1raw1of(CompletionNet):    # 1raw1of                                   model/unet.py:559
    def __init__(self, features_root=64, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True,
                 padding=True, **kw):
        super().__init__('1raw1of', features_root, tot_raw_num, tot_of_num, border_mode, rawRange, useFlow, padding, **kw)
