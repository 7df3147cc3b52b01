"""train.py / test.py flows on a synthetic UCSDped2-shaped dataset: CPU part = config, bbox / foreground stages and the
artefact names; GPU part = the whole pipeline through the root-level entry points down to a frame-level AUROC."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import _synthetic_dataset as syn
from vec_vad_b200 import pipeline as pl

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_and_foreground_stages(tmp_path, monkeypatch):
    root = syn.make(str(tmp_path / 'ws'), cfg_overrides={'context_of_num': 0, 'epochs': 1})
    monkeypatch.chdir(root)
    cfg = pl.Config('config.cfg', 'train')
    assert (cfg.tot_frame_num, cfg.tot_of_num, cfg.rawRange, cfg.padding, cfg.batch_size) == (5, 1, None, False, 128)
    probe = pl.vd.unified_dataset_interface('UCSDped2', os.path.join('raw_datasets', 'UCSDped2'), context_frame_num=1, mode='train',
                                            border_mode='hard')
    boxes = pl.load_or_make_bboxes(cfg, probe)
    assert len(boxes) == 22 and boxes[0].shape == (3, 4)
    fs, fs2 = pl.extract_foreground_train(cfg, boxes)
    assert fs[0][0].shape == (66, 5, 32, 32, 3) and fs[0][0].dtype == np.uint8            # 22 frames x 3 boxes, 1x1 blocks
    assert fs2[0][0].shape == (66, 32, 32, 2) and fs2[0][0].dtype == np.float32           # context_of_num = 0: no time axis (vad_datasets.py:132-135 adds it)
    assert os.path.exists(os.path.join('data', 'raw2flow', 'UCSDped2_foreground_train_obj_det_with_motion-raw.npy'))
    back = np.load(os.path.join('data', 'raw2flow', 'UCSDped2_foreground_train_obj_det_with_motion-flow.npy'), allow_pickle=True)
    assert np.array_equal(back[0][0], fs2[0][0])
    cfg_t = pl.Config('config.cfg', 'test')
    probe_t = pl.vd.unified_dataset_interface('UCSDped2', os.path.join('raw_datasets', 'UCSDped2'), context_frame_num=1, mode='test',
                                              border_mode='hard')
    fs, fs2, fb, scene = pl.extract_foreground_test(cfg_t, pl.load_or_make_bboxes(cfg_t, probe_t))
    assert len(fs) == 10 and fs[3][0][0].shape == (3, 5, 32, 32, 3) and fb[3][0][0].shape == (3, 4) and scene is None
    # boxes without a detector: grid patches (fore_det/simple_patch.py) in x-major order
    g = pl.get_patch_loc(240, 360, 3, 4)
    assert g.shape == (12, 4) and np.allclose(g[1], [0, 79.6666666, 90, 159.6666666]) and np.allclose(g[-1, 2:], [359, 239])
    net = pl.build_network(cfg)
    assert type(net).__name__ == 'SelfCompleteNet4' and len(net._plan_list) == 6


def test_detector_modes_need_shipped_bboxes(tmp_path, monkeypatch):
    root = syn.make(str(tmp_path / 'ws'), cfg_overrides={'train_bbox_saved': 'False'})
    monkeypatch.chdir(root)
    cfg = pl.Config('config.cfg', 'train')
    probe = pl.vd.unified_dataset_interface('UCSDped2', os.path.join('raw_datasets', 'UCSDped2'), context_frame_num=1, mode='train',
                                            border_mode='hard')
    with pytest.raises(RuntimeError, match='mmdet'):
        pl.load_or_make_bboxes(cfg, probe)


@pytest.mark.gpu
def test_train_and_test_entry_points_end_to_end(tmp_path):
    root = syn.make(str(tmp_path / 'ws'), cfg_overrides={'context_of_num': 0, 'epochs': 6, 'batch_size': 32})
    env = dict(os.environ, PYTHONPATH=REPO)
    for script in ('train.py', 'test.py'):
        r = subprocess.run([sys.executable, os.path.join(REPO, script)], cwd=root, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'AUC@ROC is' in r.stdout
    d = os.path.join(root, 'data', 'raw2flow')
    models = torch.load(os.path.join(d, 'UCSDped2_model_obj_det_with_motion_SelfComplete.npy'), weights_only=False)
    sd = models[0][0][0]
    assert all(k.startswith('module.') for k in sd) and 'module.outc_of.conv.weight' in sd          # loadable by the reference's test.py
    res = np.load(os.path.join(root, 'results', 'UCSDped2', 'raw2flow_obj_det_with_motion_SelfComplete_frame_results.npz'))
    assert float(res['roc_auc']) > 0.9          # fast squares score higher than the slow ones the UNets were trained on
    mask = torch.load(os.path.join(root, 'results', 'UCSDped2', 'score_mask', '7'), weights_only=False)
    assert mask.shape == (240, 360) and mask.min() == -100000


@pytest.mark.gpu
def test_batched_frame_scoring_equals_frame_by_frame(tmp_path, monkeypatch):
    """score_frames batches the cubes of many frames per model call; the per-frame score masks must equal the reference's
    one-forward-per-(frame, block) schedule (max_cubes=1 forces it) bit for bit."""
    root = syn.make(str(tmp_path / 'ws'), cfg_overrides={'context_of_num': 0})
    monkeypatch.chdir(root)
    cfg = pl.Config('config.cfg', 'test')
    probe = pl.vd.unified_dataset_interface('UCSDped2', os.path.join('raw_datasets', 'UCSDped2'), context_frame_num=1, mode='test',
                                            border_mode='hard')
    fs, fs2, fb, scene = pl.extract_foreground_test(cfg, pl.load_or_make_bboxes(cfg, probe))
    torch.manual_seed(0)
    net = pl.build_network(cfg).cuda().eval()
    dev = torch.device('cuda', 0)
    a = pl.score_frames(cfg, lambda s, h, w: net, lambda s, h, w: ((3.0, 2.0), (1.0, 0.5)), fs, fs2, fb, dev)
    b = pl.score_frames(cfg, lambda s, h, w: net, lambda s, h, w: ((3.0, 2.0), (1.0, 0.5)), fs, fs2, fb, dev, max_cubes=1)
    assert len(a) == len(b) == 10
    for ma, mb in zip(a, b):
        assert np.array_equal(ma, mb) and ma.max() > -100000
    # bounded workspace: sub-batches of 7 cubes through a workspace shared with a second model give the same masks
    pool = pl.vu.WorkspacePool()
    torch.manual_seed(1)
    other = pl.build_network(cfg).cuda().eval().share_workspace(pool)
    net.share_workspace(pool)
    other.score(torch.rand(3, 15, 32, 32, device='cuda'), torch.randn(3, 2, 32, 32, device='cuda'))      # overwrites the shared bytes
    c = pl.score_frames(cfg, lambda s, h, w: net, lambda s, h, w: ((3.0, 2.0), (1.0, 0.5)), fs, fs2, fb, dev, score_batch=7)
    for ma, mc in zip(a, c):
        assert np.array_equal(ma, mc)
    assert net._ws_batch <= 7 and net._parts[0]['ws'] is pool.buf
