"""FlowNet2 inference graph.

CPU : oracle/flownet2_oracle.py against tests/golden/flownet2.npz (written by the reference's own FlowNet2 class,
      tests/golden/make_flownet2_golden.py); vec_vad_b200.flownet2.FlowNet2's state_dict keys / shapes / order against the reference's.
GPU : every layer kernel of csrc/flownet_ops.cu through the C ABI against torch.nn.functional on the CPU (fp32; the kernels are exact
      fp32 FMA tiles, so the tolerance covers summation order only), then the whole stack against the reference fixture.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import flownet2_oracle as fno

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'flownet2.npz')


@pytest.fixture(scope='module')
def gold():
    g = np.load(GOLD)
    keys = [str(k) for k in g['keys']]
    shapes = [tuple(int(v) for v in s[:n]) for s, n in zip(g['shapes'], g['ndims'])]
    frames = torch.from_numpy(g['frames'].astype(np.float32)).permute(3, 0, 1, 2)[None].contiguous()      # [1,3,2,H,W]
    return dict(keys=keys, shapes=shapes, inputs=frames, flow=torch.from_numpy(g['flow']),
                inter={k[len('oracle_'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('oracle_')})


def test_oracle_matches_reference_fixture(gold):
    sd = fno.seeded_state(zip(gold['keys'], gold['shapes']))
    out, inter = fno.flownet2_forward(sd, gold['inputs'])
    # identical arithmetic, but oneDNN's summation order depends on the host's core count / thread setting: 3e-5 of the flow range
    tol = 3e-5 * float(gold['flow'].abs().max())
    assert torch.allclose(out, gold['flow'], rtol=1e-4, atol=tol), float((out - gold['flow']).abs().max())
    for k, v in gold['inter'].items():
        assert torch.allclose(inter[k], v, rtol=1e-4, atol=3e-5 * float(v.abs().max())), k


def test_state_dict_surface_equals_reference(gold):
    from vec_vad_b200.flownet2 import FlowNet2
    m = FlowNet2()
    sd = m.state_dict()
    assert list(sd.keys()) == gold['keys']
    assert [tuple(v.shape) for v in sd.values()] == gold['shapes']
    assert sum(v.numel() for v in sd.values()) == 162518834                     # FlowNet2's published parameter count
    with pytest.raises(NotImplementedError):
        FlowNet2(with_bn=True)


# ------------------------------------------------------------------------------------------------------------------ GPU
def _views():
    from vec_vad_b200 import flownet2 as fn
    return fn


# (cin, cout, k, stride, h, w, leaky) at batch 2.  Small grids, ragged sizes (scalar epilogue, part-filled tiles) and shapes large
# enough that the planner (csrc/flownet_ops.cu plan_conv) leaves the 128-pixel tiles: together they use every tile shape, which
# test_conv_plan_covers_every_tile_shape keeps true.
CONV_CASES = [(3, 64, 7, 2, 64, 96, 1), (12, 64, 7, 2, 32, 64, 1), (64, 128, 5, 2, 24, 40, 1), (37, 50, 3, 1, 19, 23, 1),
              (256, 32, 1, 1, 8, 12, 1), (130, 2, 3, 1, 16, 16, 0), (16, 16, 3, 2, 33, 31, 0), (82, 16, 3, 1, 40, 24, 0),
              (20, 16, 3, 1, 160, 192, 1), (20, 13, 3, 1, 157, 191, 0), (24, 32, 3, 1, 128, 160, 1), (40, 70, 3, 1, 61, 67, 1),
              (24, 130, 3, 1, 90, 131, 1), (32, 128, 3, 1, 64, 96, 0), (30, 60, 3, 1, 50, 77, 1)]
ALL_TILES = {(128, 16), (128, 32), (128, 64), (128, 128), (256, 64), (256, 32), (512, 16)}


def _plan(cin, h, w, cout, k, s, transposed, batch, scratch_floats):
    import ctypes as C
    from vec_vad_b200 import _lib
    tm, tn, ks = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().vecvad_fn_conv_plan(cin, h, w, cout, k, s, transposed, batch, scratch_floats, C.byref(tm), C.byref(tn), C.byref(ks)),
               'fn_conv_plan')
    return tm.value, tn.value, ks.value


def test_conv_plan_covers_every_tile_shape():
    """Host-only: the plan is a function of the shape alone, respects the scratch bound, and the GPU cases above reach every tile."""
    scratch = 16 << 20                                                          # flownet2._scratch's default: 64 MB
    used = set()
    for cin, cout, k, s, h, w, _ in CONV_CASES:
        tm, tn, ks = _plan(cin, h, w, cout, k, s, 0, 2, scratch)
        assert (tm, tn, ks) == _plan(cin, h, w, cout, k, s, 0, 2, scratch)
        assert (tm, tn) in ALL_TILES and 1 <= ks <= 32
        oh, ow = (h + 2 * ((k - 1) // 2) - k) // s + 1, (w + 2 * ((k - 1) // 2) - k) // s + 1
        assert ks == 1 or ks * cout * 2 * oh * ow <= scratch
        assert _plan(cin, h, w, cout, k, s, 0, 2, 0)[2] == 1                     # no scratch buffer, no split
        used.add((tm, tn))
    assert used == ALL_TILES, sorted(ALL_TILES - used)
    tm, tn, ks = _plan(770, 24, 32, 128, 4, 2, 1, 1, scratch)                    # a transposed conv of FlowNetS: four phases, one plan
    assert (tm, tn) in ALL_TILES and ks * 4 * 128 * 24 * 32 <= scratch
    from vec_vad_b200 import _lib
    import ctypes as C
    z = C.c_int()
    assert _lib.lib().vecvad_fn_conv_plan(8, 8, 8, 8, 4, 1, 0, 1, 0, C.byref(z), C.byref(z), C.byref(z)) != 0   # kernel size 4 is no Conv2d here


@pytest.mark.gpu
@pytest.mark.parametrize('cin,cout,k,s,h,w,leaky', CONV_CASES)
def test_conv2d_kernel_matches_torch(cin, cout, k, s, h, w, leaky):
    fn = _views()
    g = torch.Generator().manual_seed(cin * 131 + cout)
    B = 2
    src_buf = torch.randn(B, cin + 5, h, w, generator=g)                         # the input is channels [2, 2+cin) of a larger buffer
    wt, bias = torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k)) ** 0.5, torch.randn(cout, generator=g)
    want = F.conv2d(src_buf[:, 2:2 + cin].double(), wt.double(), bias.double(), s, (k - 1) // 2)
    if leaky:
        want = F.leaky_relu(want, 0.1)
    oh, ow = want.shape[2:]
    dst_buf = torch.full((B, cout + 3, oh, ow), 7.0).cuda()                       # ... and the output a slice too: neighbours untouched
    from vec_vad_b200 import _lib
    sv, dv = fn.View(src_buf.cuda(), 2, 2 + cin), fn.View(dst_buf, 1, 1 + cout)
    wc, bc = wt.cuda(), bias.cuda()
    sc = fn._scratch(wc.device)                                                   # small grids split the contraction through it
    _lib.check(_lib.lib().vecvad_fn_conv2d(sv.ptr, sv.bs, cin, h, w, _lib.ptr(wc), _lib.ptr(bc), dv.ptr, dv.bs, cout, k, s, leaky, B,
                                           _lib.ptr(sc), sc.numel(), _lib.cur_stream()), 'fn_conv2d')
    got = dst_buf.cpu()
    assert torch.allclose(got[:, 1:1 + cout].double(), want, rtol=1e-4, atol=1e-5), float((got[:, 1:1 + cout].double() - want).abs().max())
    assert bool((got[:, 0] == 7.0).all()) and bool((got[:, 1 + cout:] == 7.0).all())


@pytest.mark.gpu
@pytest.mark.parametrize('cin,cout,h,w,bias,leaky', [(1024, 512, 2, 3, True, 1), (386, 64, 16, 24, True, 1), (2, 2, 5, 7, True, 0), (2, 2, 9, 4, False, 0),
                                                     (162, 16, 20, 12, True, 1)])
def test_deconv_kernel_matches_torch(cin, cout, h, w, bias, leaky):
    fn = _views()
    g = torch.Generator().manual_seed(cin + 7 * cout)
    B = 2
    x = torch.randn(B, cin, h, w, generator=g)
    wt = torch.randn(cin, cout, 4, 4, generator=g) * (4.0 / (cin * 16)) ** 0.5
    b = torch.randn(cout, generator=g) if bias else None
    want = F.conv_transpose2d(x.double(), wt.double(), None if b is None else b.double(), 2, 1)
    if leaky:
        want = F.leaky_relu(want, 0.1)
    net = fn._SubNet('Fusion')                                                    # any table: the layer below is injected
    layer = fn._Layer('deconv' if bias else 'upflow_nobias', cin, cout, 4)
    with torch.no_grad():
        layer.weight.copy_(wt)
        if bias:
            layer.bias.copy_(b)
    layer = layer.cuda()
    net.spec['probe'] = ('deconv' if leaky else ('upflow' if bias else 'upflow_nobias'), cin, cout, 4, 2, layer)
    got = net.deconv('probe', fn.View(x.cuda())).dense().cpu()
    assert torch.allclose(got.double(), want, rtol=1e-4, atol=1e-5), float((got.double() - want).abs().max())


@pytest.mark.gpu
def test_normalize_upsample_scale_copy_match_torch():
    fn = _views()
    from vec_vad_b200 import _lib
    g = torch.Generator().manual_seed(5)
    ims = torch.rand(2, 3, 2, 64, 128, generator=g) * 255
    mean = ims.view(2, 3, -1).mean(-1).view(2, 3, 1, 1, 1)
    want = (ims - mean) / 255.0
    want = torch.cat((want[:, :, 0], want[:, :, 1]), 1)
    x = torch.empty(2, 6, 64, 128).cuda()
    scratch = torch.empty(6, dtype=torch.float64).cuda()
    ic = ims.cuda()
    _lib.check(_lib.lib().vecvad_fn_normalize_pair(_lib.ptr(ic), _lib.ptr(x), _lib.ptr(scratch), 2, 64, 128, 255.0, _lib.cur_stream()), 'normalize')
    assert torch.allclose(x.cpu(), want, rtol=1e-5, atol=2e-6)
    f = torch.randn(2, 2, 7, 9, generator=g)
    for mode in ('bilinear', 'nearest'):
        out = torch.empty(2, 5, 28, 36).cuda()
        fn.upsample4(fn.View(f.cuda()), fn.View(out, 1, 3), mode, 20.0)
        want = F.interpolate(f * 20.0, scale_factor=4, mode=mode, **({'align_corners': False} if mode == 'bilinear' else {}))
        assert torch.allclose(out[:, 1:3].cpu(), want, rtol=1e-5, atol=1e-5), mode
    src = torch.randn(3, 4, 5, 6, generator=g)
    dst = torch.zeros(3, 9, 5, 6).cuda()
    fn.scale_copy(fn.View(src.cuda(), 1, 3), fn.View(dst, 6, 8), 0.5, 0.1)
    assert torch.allclose(dst[:, 6:8].cpu(), F.leaky_relu(src[:, 1:3] * 0.5, 0.1)) and float(dst[:, :6].abs().sum()) == 0.0


@pytest.mark.gpu
def test_flownet2_stack_matches_reference_fixture(gold):
    from vec_vad_b200.flownet2 import FlowNet2
    m = FlowNet2()
    m.load_state_dict(fno.seeded_state(zip(gold['keys'], gold['shapes'])))
    m = m.cuda().eval()
    out, inter = m(gold['inputs'].cuda(), return_intermediates=True)
    rep = {}
    for k, v in gold['inter'].items():
        d = float((inter[k].cpu() - v).abs().max())
        rep[k] = (d, float(v.abs().max()))
        assert d <= 2e-4 * float(v.abs().max()) + 1e-5, (k, rep)
    d = float((out.cpu() - gold['flow']).abs().max())
    assert d <= 2e-4 * float(gold['flow'].abs().max()), (d, rep)
    # batch of two identical pairs: every image of the batch gets the same flow (no cross-image coupling in any kernel)
    out2 = m(torch.cat([gold['inputs'], gold['inputs']]).cuda())
    assert torch.allclose(out2[0], out2[1], atol=1e-6) and torch.allclose(out2[0].cpu(), gold['flow'][0], atol=2e-4 * float(gold['flow'].abs().max()))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 2, 100, 128).cuda())
    with pytest.raises(RuntimeError):
        m(gold['inputs'])                                                          # CPU tensor: no fallback


@pytest.mark.gpu
def test_calc_optical_flow_stage_writes_the_flow_modality(tmp_path, monkeypatch):
    """calc_optical_flow.py:12-88 on a tiny on-disk dataset: one .npy per frame in the raw_datasets layout, frame-sized [H,W,2] float32,
    computed from the first two entries of the clamped window at a video's borders and from (current, next) elsewhere."""
    import cv2
    from tests import _synthetic_dataset as syn
    from vec_vad_b200 import optical_flow as of, vad_datasets as vd
    from vec_vad_b200.flownet2 import FlowNet2
    root = syn.make(str(tmp_path / 'ws'), n_train=(3,), n_test=(2,))
    monkeypatch.chdir(root)
    torch.manual_seed(1)
    net = FlowNet2().cuda().eval()
    ds = vd.unified_dataset_interface(dataset_name='UCSDped2', dir=os.path.join('raw_datasets', 'UCSDped2'), context_frame_num=1, mode='train',
                                      border_mode='hard')
    of.calc_optical_flow(ds, net=net, of_root_dir='./of_out', verbose=False)
    frames = [cv2.imread(a) for a in ds.all_frame_addr]
    assert len(frames) == 3
    # windows [0,0,1] / [0,1,2] / [1,2,2]: a border window feeds its FIRST two entries (frame 0 against itself at the very start,
    # calc_optical_flow.py:44-56), every other window its last two
    for idx, pair in ((0, (0, 0)), (1, (1, 2)), (2, (1, 2))):
        saved = np.load(os.path.join('of_out', 'UCSDped2', 'Train', 'Train001', '%03d.npy' % (idx + 1)))
        assert saved.shape == (240, 360, 2) and saved.dtype == np.float32
        ims = np.array([[cv2.resize(frames[pair[0]], (512, 384)), cv2.resize(frames[pair[1]], (512, 384))]]).transpose((0, 4, 1, 2, 3))
        want = cv2.resize(net(torch.from_numpy(ims.astype(np.float32)).cuda())[0].cpu().numpy().transpose((1, 2, 0)), (360, 240))
        assert np.allclose(saved, want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()))


def test_calc_optical_flow_batched_writes_the_same_files(tmp_path, monkeypatch):
    """Host logic of the optical-flow stage (calc_optical_flow.py:12-88) with a stand-in network on the CPU: grouping ``batch_pairs`` frame
    pairs per network call writes the same files, in the same layout, as one pair per call -- including a ragged last group."""
    import cv2
    from tests import _synthetic_dataset as syn
    from vec_vad_b200 import optical_flow as of, vad_datasets as vd
    root = syn.make(str(tmp_path / 'ws'), n_train=(3, 2), n_test=(2,))
    monkeypatch.chdir(root)

    class PerImageNet(torch.nn.Module):                        # [B,3,2,384,512] -> [B,2,384,512], every image on its own
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor([0.5, -0.25]))
            self.calls = []

        def forward(self, ims):
            self.calls.append(ims.shape[0])
            assert tuple(ims.shape[1:]) == (3, 2, 384, 512) and ims.dtype == torch.float32
            d = (ims[:, :, 1] - ims[:, :, 0]).mean(1, keepdim=True) + ims[:, :, 0].mean(1, keepdim=True) * 0.01
            return (d * self.w.view(1, 2, 1, 1)).detach()

    ds = vd.unified_dataset_interface(dataset_name='UCSDped2', dir=os.path.join('raw_datasets', 'UCSDped2'), context_frame_num=1, mode='train',
                                      border_mode='hard')
    one, four = PerImageNet(), PerImageNet()
    of.calc_optical_flow(ds, net=one, of_root_dir='./of_one', verbose=False)
    of.calc_optical_flow(ds, net=four, of_root_dir='./of_four', verbose=False, batch_pairs=4)
    assert one.calls == [1] * 5 and four.calls == [4, 1]
    files = sorted(os.path.relpath(os.path.join(d, f), 'of_one') for d, _, fs in os.walk('of_one') for f in fs)
    assert len(files) == 5 and files == sorted(os.path.relpath(os.path.join(d, f), 'of_four') for d, _, fs in os.walk('of_four') for f in fs)
    frame = cv2.imread(ds.all_frame_addr[0])
    for f in files:
        a, b = np.load(os.path.join('of_one', f)), np.load(os.path.join('of_four', f))
        assert a.shape == frame.shape[:2] + (2,) and a.dtype == np.float32 and np.array_equal(a, b) and np.abs(a).max() > 0
