"""get_foreground's crop + resize: an integer path -- bit-exact.

CPU: oracle/resize_oracle.py (numpy restatement of cv2.resize INTER_LINEAR for uint8 / float32) against the committed fixture
     tests/golden/crop_resize.npz that the REFERENCE's get_foreground wrote (tests/golden/make_crop_resize_golden.py), and
     against cv2 itself on fresh random crops (cv2 is the third-party dependency the arithmetic lives in).
GPU: vec_vad_b200.vad_datasets.get_foreground_device (csrc/crop_resize.cu through the C ABI) against the same fixture, against the
     oracle on fresh boxes at full UCSDped2 frame size, both frame layouts, and the dataset hook (__getitem__ with
     foreground_device set) against the host path.
"""
import os

import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro
from vec_vad_b200 import vad_datasets as vd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crop_resize.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def _random_boxes(rng, H, W, n):
    x0, y0 = rng.uniform(0, W - 2, n), rng.uniform(0, H - 2, n)
    x1, y1 = rng.uniform(x0 + 1, W - 0.01), rng.uniform(y0 + 1, H - 0.01)
    return np.stack([x0, y0, x1, y1], 1).astype(np.float32)


def test_oracle_matches_reference_fixture(gold):
    assert np.array_equal(ro.get_foreground(gold['raw'], gold['boxes'], 32), gold['raw_patches'])
    assert np.array_equal(ro.get_foreground(gold['flow'], gold['boxes'], 32), gold['flow_patches'])
    assert np.array_equal(ro.get_foreground(gold['raw'][0], gold['boxes'], 32), gold['raw3_patches'])
    assert np.array_equal(ro.get_foreground(gold['raw'], gold['boxes'][:10], 16), gold['raw_patches_p16'])


def test_oracle_matches_cv2_on_fresh_crops():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(5)
    sizes = [(64, 64), (32, 32), (1, 1), (2, 5), (300, 200), (33, 31), (64, 32), (32, 64), (17, 90), (128, 128), (96, 96), (65, 64), (5, 5)]
    sizes += [(int(rng.integers(1, 240)), int(rng.integers(1, 360))) for _ in range(120)]
    for (h, w) in sizes:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(a, (32, 32)), ro.resize_linear(a, 32)), (h, w, 'uint8')
        f = (rng.standard_normal((h, w, 2)) * 3).astype(np.float32)
        assert np.array_equal(cv2.resize(f, (32, 32)), ro.resize_linear(f, 32)), (h, w, 'float32')


def test_host_get_foreground_matches_fixture(gold):
    assert np.array_equal(vd.get_foreground(gold['raw'], gold['boxes'], 32), gold['raw_patches'])
    assert np.array_equal(vd.get_foreground(gold['flow'], gold['boxes'], 32), gold['flow_patches'])


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_device_get_foreground_matches_reference_fixture(gold):
    raw, flow, boxes = torch.from_numpy(gold['raw']).cuda(), torch.from_numpy(gold['flow']).cuda(), gold['boxes']
    assert np.array_equal(vd.get_foreground_device(raw, boxes, 32).cpu().numpy(), gold['raw_patches'])
    assert np.array_equal(vd.get_foreground_device(flow, boxes, 32).cpu().numpy(), gold['flow_patches'])
    assert np.array_equal(vd.get_foreground_device(raw[0], boxes, 32).cpu().numpy(), gold['raw3_patches'])
    assert np.array_equal(vd.get_foreground_device(raw, boxes[:10], 16).cpu().numpy(), gold['raw_patches_p16'])
    # cv2's own frame layout ([T,H,W,C]) without a host transpose: same patches
    hwc = torch.from_numpy(np.ascontiguousarray(np.transpose(gold['raw'], [0, 2, 3, 1]))).cuda()
    assert np.array_equal(vd.get_foreground_device(hwc, boxes, 32, layout='HWC').cpu().numpy(), gold['raw_patches'])


@pytest.mark.gpu
def test_device_get_foreground_matches_oracle_at_frame_size():
    rng = np.random.default_rng(77)
    H, W = 240, 360                                           # UCSDped2 frames, 5-frame context stack
    raw = rng.integers(0, 256, (5, 3, H, W), dtype=np.uint8)
    flow = (rng.standard_normal((5, 2, H, W)) * 4).astype(np.float32)
    boxes = _random_boxes(rng, H, W, 200)
    want_raw, want_flow = ro.get_foreground(raw, boxes, 32), ro.get_foreground(flow, boxes, 32)
    assert np.array_equal(vd.get_foreground_device(torch.from_numpy(raw).cuda(), boxes, 32).cpu().numpy(), want_raw)
    assert np.array_equal(vd.get_foreground_device(torch.from_numpy(flow).cuda(), boxes, 32).cpu().numpy(), want_flow)
    # no boxes: an empty tensor of the right shape, no launch
    assert tuple(vd.get_foreground_device(torch.from_numpy(raw).cuda(), boxes[:0], 32).shape) == (0, 5, 3, 32, 32)
    with pytest.raises(ValueError):
        vd.get_foreground_device(torch.from_numpy(raw).cuda(), np.array([[10.2, 10.2, 10.4, 30.0]]), 32)     # empty after ceil
    with pytest.raises(RuntimeError):
        vd.get_foreground_device(torch.from_numpy(raw), boxes, 32)                                            # CPU tensor: no fallback


@pytest.mark.gpu
def test_dataset_hook_is_bit_identical_to_host_path(tmp_path):
    from tests import _synthetic_dataset as sd
    root = sd.make(str(tmp_path))
    ddir = os.path.join(root, 'raw_datasets', 'UCSDped2')
    bboxes = np.load(os.path.join(ddir, 'bboxes_train_obj_det_with_motion.npy'), allow_pickle=True)
    for sub, fmt, ctx in (('raw_datasets', '.tif', 4), ('optical_flow', '.npy', 0)):
        kw = dict(dataset_name='UCSDped2', dir=os.path.join(root, sub, 'UCSDped2'), mode='train', context_frame_num=ctx, border_mode='predict',
                  all_bboxes=bboxes, patch_size=32, file_format=fmt)
        host = vd.unified_dataset_interface(**kw)
        dev = vd.unified_dataset_interface(**kw)
        dev.foreground_device = torch.device('cuda')
        n = 0
        for i in range(len(host)):
            if len(bboxes[i]) == 0:
                continue
            a, b = host[i][0], dev[i][0]
            assert b.is_cuda and a.dtype == b.dtype and torch.equal(a, b.cpu())
            n += 1
        assert n > 0
