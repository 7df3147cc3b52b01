"""CPU checks of the dataset surface against fixtures dumped from the real reference (tests/golden/dataset.npz,
index_paths.npz): cube_to_train_dataset items, get_foreground crops, calc_block_idx, context_range, AUROC helper.
Integer / index paths are compared bit-exactly."""
import os

import numpy as np
import pytest
import torch

from vec_vad_b200 import utils as vutils
from vec_vad_b200 import vad_datasets as vd


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(os.path.join(golden_dir, 'dataset.npz'), allow_pickle=False)


@pytest.fixture(scope='module')
def gi(golden_dir):
    return np.load(os.path.join(golden_dir, 'index_paths.npz'), allow_pickle=False)


def test_cube_to_train_dataset_items_bit_exact(g):
    ds = vd.cube_to_train_dataset(g['raw_u8'], target=g['flow'])
    assert len(ds) == 3
    a, b, c = ds[1]
    assert a.dtype == torch.float32 and np.array_equal(a.numpy(), g['item1_in'])
    assert np.array_equal(b.numpy(), g['item1_tgt']) and np.array_equal(c.numpy(), g['item1_copy'])
    ds4 = vd.cube_to_train_dataset(g['raw_u8'][:, 0], target=g['flow'][:, 0])      # 4-D inputs get a T=1 axis
    a4, b4, _ = ds4[2]
    assert np.array_equal(a4.numpy(), g['item4_in']) and np.array_equal(b4.numpy(), g['item4_tgt'])
    # default DataLoader collate gives the batch layout the engine expects
    from torch.utils.data import DataLoader
    x, x_of, _ = next(iter(DataLoader(ds, batch_size=3, shuffle=False)))
    assert tuple(x.shape) == (3, 15, 32, 32) and tuple(x_of.shape) == (3, 10, 32, 32)


def test_get_foreground_bit_exact(g):
    assert np.array_equal(vd.get_foreground(g['img'], g['boxes'], 32), g['fg4'])
    assert np.array_equal(vd.get_foreground(g['img'][0], g['boxes'], 32), g['fg3'])
    assert np.array_equal(vd.get_foreground(g['imgf'], g['boxes'], 32), g['fgf'])


@pytest.mark.parametrize('mode', [1, 5, 9])
def test_calc_block_idx_bit_exact(g, mode):
    for r, want in zip(g['block_boxes'], g['block_idx_mode%d' % mode]):
        x0, x1 = sorted([r[0] * 360, r[2] * 360])
        y0, y1 = sorted([r[1] * 240, r[3] * 240])
        got = sorted(vutils.calc_block_idx(x0, x1, y0, y1, 240 / 3, 360 / 4, mode=mode))
        flat = -np.ones(18, dtype=np.int64)
        flat[:2 * len(got)] = np.array(got).reshape(-1)
        assert np.array_equal(flat, want)


def test_context_range_bit_exact(gi):
    n_cases = 0
    for key in gi.files:
        if '|' not in key:
            continue
        lname, mode, c = key.split('|')
        lens = gi['layout_' + lname]
        fvi = []
        for v, n in enumerate(lens):
            fvi += [v + 1] * int(n)
        want = gi[key]
        for i in range(len(fvi)):
            try:
                got = np.array(vd.context_range(i, fvi, int(c), mode))
            except NotImplementedError:
                got = -np.ones(want.shape[1], dtype=np.int64)
            assert np.array_equal(got, want[i]), (key, i, got, want[i])
            n_cases += 1
    assert n_cases > 500


def test_auroc_helper_matches_reference(gi, tmp_path):
    auc = vutils.save_roc_pr_curve_data(gi['roc_scores'], gi['roc_labels'], str(tmp_path / 'r.npz'), verbose=False)
    assert auc == float(gi['roc_auc'])
    assert set(np.load(str(tmp_path / 'r.npz')).files) >= {'preds', 'truth', 'fpr', 'tpr', 'roc_auc', 'pr_auc_norm', 'pr_auc_anom'}


def test_paint_score_mask_is_running_max_of_ceiled_rectangles():
    h, w, big = 24, 36, 100000
    rng = np.random.RandomState(3)
    boxes = np.stack([rng.uniform(0, 20, 8), rng.uniform(0, 12, 8), rng.uniform(20, 35.5, 8), rng.uniform(12, 23.5, 8)], 1)
    scores = rng.randn(8)
    want = -1 * np.ones((h, w)) * big
    for m in range(8):                       # the literal sequence of test.py:350-357
        mask = -1 * np.ones((h, w)) * big
        x_min, x_max = int(np.ceil(boxes[m][0])), int(np.ceil(boxes[m][2]))
        y_min, y_max = int(np.ceil(boxes[m][1])), int(np.ceil(boxes[m][3]))
        mask[y_min:y_max, x_min:x_max] = scores[m]
        want = np.max(np.concatenate([want[:, :, None], mask[:, :, None]], axis=2), axis=2)
    got = vutils.paint_score_mask(-1 * np.ones((h, w)) * big, scores, boxes, big)
    assert np.array_equal(got, want)


def test_frame_folder_datasets(tmp_path):
    """Directory scan + context assembly on a synthetic UCSDped2-shaped folder (two train videos)."""
    import cv2
    root = tmp_path / 'UCSDped2'
    for v, n in (('Train001', 6), ('Train002', 5)):
        d = root / 'Train' / v
        d.mkdir(parents=True)
        for i in range(n):
            cv2.imwrite(str(d / ('%03d.tif' % (i + 1))), np.full((24, 36, 3), 10 * (i + 1) + (100 if v.endswith('2') else 0), np.uint8))
    ds = vd.unified_dataset_interface('UCSDped2', str(root), mode='train', context_frame_num=4, border_mode='predict')
    assert len(ds) == 11 and ds.frame_video_idx == [1] * 6 + [2] * 5
    img, z = ds[7]                               # second frame of video 2: padded with the video's first frame
    assert tuple(img.shape) == (5, 3, 24, 36) and float(z) == 0.0
    assert [int(img[t, 0, 0, 0]) for t in range(5)] == [110, 110, 110, 110, 120]
    boxes = [np.array([[2.2, 3.1, 20.0, 19.5]], dtype=np.float32)] * 11
    ds2 = vd.unified_dataset_interface('UCSDped2', str(root), mode='train', context_frame_num=4, border_mode='predict', all_bboxes=boxes,
                                       patch_size=32)
    cube, _ = ds2[3]
    assert tuple(cube.shape) == (1, 5, 3, 32, 32)
    with pytest.raises(NotImplementedError):
        vd.unified_dataset_interface('nope', str(root))
