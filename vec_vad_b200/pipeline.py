"""The reference's ``train.py`` / ``test.py`` flows as functions (the root-level scripts of the same names call them).

Same ``config.cfg`` (train.py:19-42, test.py:18-41), same on-disk artefacts and file names (bboxes, foreground cube sets,
``*_model_*.npy`` nested lists of state_dicts with the DataParallel ``module.`` prefix, training score sets, per-frame
score masks, ROC npz), same stage caching flags.  What differs is HOW the hot loops run:
  * train.py:378-402  ->  ``CompletionNet.train_step`` (forward + MSE + backward + Adam fused on the device, no per-step
    host sync; the running means printed every 5 batches are read back only when printed), cubes resident in HBM as
    uint8 (``DeviceCubeStore``) instead of a per-item Python ``DataLoader``; with WORLD_SIZE > 1 the batch is sharded
    over ranks with one NCCL gradient all-reduce per step (``ddp.GradReducer``) instead of ``nn.DataParallel``.
  * train.py:412-431 / test.py:270-345  ->  ``CompletionNet.score`` (eval-mode forward with the per-cube sum of squared
    error reduced on the device).
Foreground localisation by the Cascade R-CNN detector (fore_det/, mmdet 1.0rc0) is not part of the hot path: like the
reference's released setup (README.md:51) the shipped ``bboxes_*.npy`` are loaded; 'simple_patch' and 'frame' boxes,
which need no detector, are generated.
"""
import itertools
import os
from configparser import ConfigParser

import numpy as np
import torch

from . import ddp
from . import unet as vu
from . import vad_datasets as vd
from .utils import calc_block_idx, paint_score_mask, save_roc_pr_curve_data

BIG_NUMBER = 100000


class AverageMeter:
    """Running mean of the printed losses (helper/misc.py:59-76)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class Config:
    """Every knob the two scripts read from config.cfg, under the reference's variable names."""

    def __init__(self, path='config.cfg', split='train'):
        cp = ConfigParser()
        if not cp.read(path):
            raise FileNotFoundError(path)
        self.cp, self.split = cp, split
        g = cp.get
        self.dataset_name = g('shared_parameters', 'dataset_name')
        self.raw_dataset_dir = g('shared_parameters', 'raw_dataset_dir')
        self.foreground_extraction_mode = g('shared_parameters', 'foreground_extraction_mode')
        self.data_root_dir = g('shared_parameters', 'data_root_dir')
        self.modality = g('shared_parameters', 'modality')
        self.method = g('shared_parameters', 'method')
        self.mode = g('%s_parameters' % split, 'mode')
        d = self.dataset_name
        try:
            self.patch_size = cp.getint(d, 'patch_size')
            self.block_mode = cp.getint(d, '%s_block_mode' % split)
            self.motionThr = cp.getfloat(d, 'motionThr')
            self.h_block, self.w_block = cp.getint(d, 'h_block'), cp.getint(d, 'w_block')
            self.bbox_saved = cp.getboolean(d, '%s_bbox_saved' % split)
            self.foreground_saved = cp.getboolean(d, '%s_foreground_saved' % split)
            self.scores_saved = cp.getboolean(d, 'scores_saved')
        except Exception:
            raise NotImplementedError
        if self.method != 'SelfComplete':
            raise NotImplementedError
        m = self.method
        self.epochs, self.batch_size = cp.getint(m, 'epochs'), cp.getint(m, 'batch_size')
        self.useFlow, self.border_mode = cp.getboolean(m, 'useFlow'), g(m, 'border_mode')
        self.context_frame_num, self.context_of_num = cp.getint(m, 'context_frame_num'), cp.getint(m, 'context_of_num')
        if self.border_mode == 'predict':
            self.tot_frame_num, self.tot_of_num = self.context_frame_num + 1, self.context_of_num + 1
        else:
            self.tot_frame_num, self.tot_of_num = 2 * self.context_frame_num + 1, 2 * self.context_of_num + 1
        self.rawRange = cp.getint(m, 'rawRange')
        if self.rawRange >= self.tot_frame_num:           # out of range = all frames (train.py:252-254)
            self.rawRange = None
        self.padding = cp.getboolean(m, 'padding')
        self.nf = cp.getint(m, 'nf')
        self.lambda_raw, self.lambda_of = cp.getfloat(m, 'lambda_raw'), cp.getfloat(m, 'lambda_of')
        self.w_raw, self.w_of = cp.getfloat(m, 'w_raw'), cp.getfloat(m, 'w_of')
        self.saveSegNum = cp.getint(d, 'saveSegNum') if cp.has_option(d, 'saveSegNum') else None
        assert self.modality == 'raw2flow'                 # train.py:258

    def path(self, name):
        return os.path.join(self.data_root_dir, self.modality, self.dataset_name + '_' + name)

    def tag(self):
        return '{}_{}'.format(self.foreground_extraction_mode, self.method)


def build_network(cfg, **kw):
    """train.py:260-268 / test.py:216-224"""
    args = dict(features_root=cfg.nf, tot_raw_num=cfg.tot_frame_num, tot_of_num=cfg.tot_of_num, border_mode=cfg.border_mode,
                rawRange=cfg.rawRange, useFlow=cfg.useFlow, padding=cfg.padding, patch_size=cfg.patch_size)
    args.update(kw)
    assert cfg.tot_frame_num == 5
    if cfg.tot_of_num == 1:
        return vu.SelfCompleteNet4(**args)
    if cfg.tot_of_num == 5:
        return vu.SelfCompleteNetFull(**args)
    raise NotImplementedError


# ------------------------------------------------------------------------------------------ stage 1: bounding boxes
def get_patch_loc(h, w, h_num, w_num):
    """Regular grid of boxes (fore_det/simple_patch.py:5-16), x-major order."""
    h_step, w_step = h / h_num, w / w_num
    ys = np.linspace(0, h - 1, h_num, endpoint=False)
    xs = np.linspace(0, w - 1, w_num, endpoint=False)
    return np.array([np.array([x, y, np.minimum(x + w_step, w - 1), np.minimum(y + h_step, h - 1)])
                     for x, y in itertools.product(tuple(xs), tuple(ys))])


def load_or_make_bboxes(cfg, dataset):
    fname = os.path.join(dataset.dir, 'bboxes_{}_{}.npy'.format(cfg.mode, cfg.foreground_extraction_mode))
    if cfg.bbox_saved:
        boxes = np.load(fname, allow_pickle=True)
        print('bboxes for {} data loaded!'.format('training' if cfg.mode == 'train' else 'testing'))
        return boxes
    h, w = vd.frame_size[cfg.dataset_name][:2]
    if cfg.foreground_extraction_mode == 'simple_patch':
        one = np.concatenate([get_patch_loc(h, w, hn, wn) for hn, wn in [(3, 4), (6, 8)]], axis=0)
    elif cfg.foreground_extraction_mode == 'frame':
        one = np.array([[0, 0, w, h]])
    elif cfg.foreground_extraction_mode in ('obj_det', 'obj_det_with_motion'):
        raise RuntimeError('foreground localisation with the Cascade R-CNN detector (fore_det/, mmdet 1.0rc0) is outside this '
                           'package: use the shipped %s (set %s_bbox_saved = True)' % (fname, cfg.split))
    else:
        raise NotImplementedError
    boxes = np.empty(len(dataset), dtype=object)
    for i in range(len(dataset)):
        boxes[i] = one
    np.save(fname, boxes)
    return boxes


# ------------------------------------------------------------------------------------------ stage 2: foreground cubes
def _motion_energy(flow_cubes):
    """Per-box motion energy that gates a cube (train.py:167-175): sum of squared flow, averaged over frames."""
    if flow_cubes.ndim == 4:
        return np.sum(flow_cubes ** 2, axis=(1, 2, 3))
    return np.mean(np.sum(flow_cubes ** 2, axis=(2, 3, 4)), axis=1)


def _cube_datasets(cfg, all_bboxes):
    fmt = vd.frame_size[cfg.dataset_name][2]
    ds = vd.unified_dataset_interface(cfg.dataset_name, os.path.join('raw_datasets', cfg.dataset_name), context_frame_num=cfg.context_frame_num,
                                      mode=cfg.mode, border_mode=cfg.border_mode, all_bboxes=all_bboxes, patch_size=cfg.patch_size,
                                      file_format=fmt)
    ds2 = vd.unified_dataset_interface(cfg.dataset_name, os.path.join('optical_flow', cfg.dataset_name),
                                       context_frame_num=cfg.context_of_num, mode=cfg.mode, border_mode=cfg.border_mode,
                                       all_bboxes=all_bboxes, patch_size=cfg.patch_size, file_format='.npy')
    # crop + resize of the foreground boxes on the GPU (bit-identical to cv2.resize, csrc/crop_resize.cu); VECVAD_DEVICE_FOREGROUND=0
    # keeps the host path
    if torch.cuda.is_available() and os.environ.get('VECVAD_DEVICE_FOREGROUND', '1') != '0':
        ds.foreground_device = ds2.foreground_device = torch.device('cuda', torch.cuda.current_device())
    return ds, ds2


def _frame_cubes(cfg, ds, ds2, idx, all_bboxes):
    """-> [(raw cube, flow cube, bbox, [(h_block, w_block), ...])] of frame idx that pass the motion gate"""
    boxes = all_bboxes[idx]
    if len(boxes) == 0:
        return []
    raw = vd.img_batch_tensor2numpy(ds[idx][0])
    flow = vd.img_batch_tensor2numpy(ds2[idx][0])
    mag = _motion_energy(flow)
    h_step, w_step = vd.frame_size[cfg.dataset_name][0] / cfg.h_block, vd.frame_size[cfg.dataset_name][1] / cfg.w_block
    out = []
    for k in range(boxes.shape[0]):
        if mag[k] > cfg.motionThr:
            blocks = calc_block_idx(boxes[k, 0], boxes[k, 2], boxes[k, 1], boxes[k, 3], h_step, w_step, mode=cfg.block_mode)
            out.append((raw[k], flow[k], boxes[k], blocks))
    return out


def _obj(nested):
    """np.save-able object array of a nested list of (ragged) arrays, as the reference writes them."""
    arr = np.empty(len(nested), dtype=object)
    for i, v in enumerate(nested):
        arr[i] = _obj(v) if isinstance(v, list) else v
    return arr


def extract_foreground_train(cfg, all_bboxes):
    """train.py:103-225.  -> (foreground_set, foreground_set2) nested [h][w] (ShanghaiTech: segment files only, returns None)."""
    ds, ds2 = _cube_datasets(cfg, all_bboxes)
    sh = cfg.dataset_name == 'ShanghaiTech'
    scenes = ds.scene_num if sh else 1

    def empty():
        return [[[[] for _ in range(cfg.w_block)] for _ in range(cfg.h_block)] for _ in range(scenes)]

    def pack(s):
        return [[[np.array(s[ss][hh][ww]) for ww in range(cfg.w_block)] for hh in range(cfg.h_block)] for ss in range(scenes)]
    fs, fs2 = empty(), empty()
    order = np.random.permutation(len(ds)) if sh else np.arange(len(ds))
    count = seg = 0
    os.makedirs(os.path.join(cfg.data_root_dir, cfg.modality), exist_ok=True)
    for ii, idx in enumerate(order):
        print('Extracting foreground in {}-th batch, {} in total'.format(ii + 1, len(ds)))
        scene = ds.scene_idx[idx] - 1 if sh else 0
        for raw, flow, _, blocks in _frame_cubes(cfg, ds, ds2, idx, all_bboxes):
            for (hb, wb) in blocks:
                fs[scene][hb][wb].append(raw)
                fs2[scene][hb][wb].append(flow)
        count += 1
        if sh and (count == cfg.saveSegNum or ii == len(ds) - 1) and count > 0:
            np.save(cfg.path('foreground_train_{}_seg_{}-raw.npy'.format(cfg.foreground_extraction_mode, seg)), _obj(pack(fs)))
            np.save(cfg.path('foreground_train_{}_seg_{}-flow.npy'.format(cfg.foreground_extraction_mode, seg)), _obj(pack(fs2)))
            fs, fs2, count, seg = empty(), empty(), 0, seg + 1
    if sh:
        print('foreground for training data saved!')
        return None, None
    a, b = pack(fs)[0], pack(fs2)[0]
    np.save(cfg.path('foreground_train_{}-raw.npy'.format(cfg.foreground_extraction_mode)), _obj(a))
    np.save(cfg.path('foreground_train_{}-flow.npy'.format(cfg.foreground_extraction_mode)), _obj(b))
    print('foreground for training data saved!')
    return a, b


def extract_foreground_test(cfg, all_bboxes):
    """test.py:101-177.  -> (foreground_set, foreground_set2, foreground_bbox_set) nested [frame][h][w], scene_idx"""
    ds, ds2 = _cube_datasets(cfg, all_bboxes)
    os.makedirs(os.path.join(cfg.data_root_dir, cfg.modality), exist_ok=True)
    scene_idx = None
    if cfg.dataset_name == 'ShanghaiTech':
        scene_idx = ds.scene_idx
        np.save(cfg.path('scene_idx.npy'), scene_idx)
    n = len(ds)
    sets = [[[[[] for _ in range(cfg.w_block)] for _ in range(cfg.h_block)] for _ in range(n)] for _ in range(3)]
    for idx in range(n):
        print('Extracting foreground in {}-th batch, {} in total'.format(idx + 1, n))
        for raw, flow, box, blocks in _frame_cubes(cfg, ds, ds2, idx, all_bboxes):
            for (hb, wb) in blocks:
                sets[0][idx][hb][wb].append(raw)
                sets[1][idx][hb][wb].append(flow)
                sets[2][idx][hb][wb].append(box)
    packed = [[[[np.array(s[ii][hh][ww]) for ww in range(cfg.w_block)] for hh in range(cfg.h_block)] for ii in range(n)] for s in sets]
    m = cfg.foreground_extraction_mode
    np.save(cfg.path('foreground_test_{}-raw.npy'.format(m)), _obj(packed[0]))
    np.save(cfg.path('foreground_test_{}-flow.npy'.format(m)), _obj(packed[1]))
    np.save(cfg.path('foreground_bbox_test_{}.npy'.format(m)), _obj(packed[2]))
    print('foreground for testing data saved!')
    return packed[0], packed[1], packed[2], scene_idx


# ------------------------------------------------------------------------------------------ stage 3: training
def _state_dict_for_disk(net):
    """CPU copy with the ``module.`` prefix the reference's DataParallel wrapper adds (train.py:410 -> test.py:256)."""
    return {'module.' + k: v.detach().cpu().clone() for k, v in net.state_dict().items()}


def load_block_state(net, sd):
    """Accepts checkpoints written by either implementation (with or without the ``module.`` prefix)."""
    if len(sd) and all(k.startswith('module.') for k in sd):
        sd = {k[7:]: v for k, v in sd.items()}
    net.load_state_dict(sd)


def train_block(net, cube_batches, cfg, reducer, meters, tag):
    """One model on one block: the hot loop of train.py:378-408.  ``cube_batches(epoch)`` yields ``(x, x_of, local_n, global_n)``
    for every GLOBAL batch (``DeviceCubeStore.rank_batches``); with several ranks every rank sees the same number of steps and
    ``x`` is None where its share of a ragged last batch is empty."""
    net.train()
    net.init_adam(lr=1e-3, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.0)       # optim.Adam(eps=1e-7, weight_decay=0.0), train.py:376
    raw_losses, of_losses = meters
    pending = []                                                                   # (device losses, batch size) since the last print
    for epoch in range(cfg.epochs):
        n_batches = 0
        for idx, (x, x_of, local_n, global_n) in enumerate(cube_batches(epoch)):
            if reducer is not None:
                reducer.set_batch(local_n, global_n)
            if x is None:
                losses = net.train_step_empty(reducer)
            else:
                losses = net.train_step(x, x_of, cfg.lambda_raw, cfg.lambda_of, reduce_grads=reducer)
            pending.append((losses, local_n))
            n_batches += 1
            if idx % 5 == 0:
                for l, n in pending:
                    lr_, lo_ = l.tolist()
                    raw_losses.update(lr_, n)
                    of_losses.update(lo_ if cfg.useFlow else 0., n)
                pending = []
                print('Block: {}, epoch {}, batch {}, raw loss: {}, of loss: {}'.format(tag, epoch, idx, raw_losses.avg, of_losses.avg))
    for l, n in pending:
        lr_, lo_ = l.tolist()
        raw_losses.update(lr_, n)
        of_losses.update(lo_ if cfg.useFlow else 0., n)


@torch.no_grad()
def score_block(net, store, batch_size, useFlow=True):
    """Per-cube sum of squared error over a cube store in order (train.py:412-431).  An empty store (a ShanghaiTech disk segment
    without cubes for this block) yields empty score vectors instead of failing the concatenate."""
    net.eval()
    raw, of = [], []
    for x, x_of in store.batches(batch_size, shuffle=False):
        r, o = net.score(x, x_of)
        raw.append(r.cpu().numpy())
        if o is not None:
            of.append(o.cpu().numpy())
    if not raw:
        return np.zeros((0,), dtype=np.float32), (np.zeros((0,), dtype=np.float32) if useFlow else [])
    return np.concatenate(raw, 0), (np.concatenate(of, 0) if of else [])


def _epoch_batches(store, cfg, rank, world):
    """One epoch of global batches of ``cfg.batch_size`` cubes, this rank's share of each (all ranks: same step count)."""
    return store.rank_batches(cfg.batch_size, rank, world, seed=ddp.shared_seed(), shuffle=True)


def default_precision():
    """Operand mode of the conv tiles for the entry points: ``VECVAD_PRECISION`` = f16 (default: fp16 operands, fp32 accumulation -- the mode
    bench.py measures; losses within 1e-5 relative of the reference arithmetic, AUROC parity in tests/test_auroc_parity_gpu.py), tf32, or
    fp32 (exact SIMT tiles)."""
    v = os.environ.get('VECVAD_PRECISION', 'f16').lower()
    if v not in ('f16', 'fp16', 'tf32', 'fp32', 'simt'):
        raise ValueError('VECVAD_PRECISION must be f16, tf32 or fp32, got %r' % v)
    return v


def train(cfg_path='config.cfg', use_tensor_cores=None):
    if use_tensor_cores is None:
        use_tensor_cores = default_precision()
    cfg = Config(cfg_path, 'train')
    rank, local, world = ddp.init_from_env()
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    probe = vd.unified_dataset_interface(cfg.dataset_name, os.path.join(cfg.raw_dataset_dir, cfg.dataset_name), context_frame_num=1,
                                         mode=cfg.mode, border_mode='hard')
    # host-side preprocessing (box files, foreground cube extraction: minutes to hours of disk-bound work on real datasets) runs
    # on rank 0 only; the others wait on a CPU (gloo) barrier with a day-long timeout instead of the NCCL watchdog's default
    slow_group = None
    if world > 1:
        import datetime
        slow_group = torch.distributed.new_group(backend='gloo', timeout=datetime.timedelta(hours=24))
    if rank == 0 or cfg.bbox_saved:
        all_bboxes = load_or_make_bboxes(cfg, probe)
    if world > 1 and not cfg.bbox_saved:
        torch.distributed.barrier(group=slow_group)
        if rank != 0:
            all_bboxes = np.load(os.path.join(probe.dir, 'bboxes_{}_{}.npy'.format(cfg.mode, cfg.foreground_extraction_mode)), allow_pickle=True)
    sh = cfg.dataset_name == 'ShanghaiTech'
    m = cfg.foreground_extraction_mode
    if not cfg.foreground_saved:
        if rank == 0:
            extract_foreground_train(cfg, all_bboxes)
        if world > 1:
            torch.distributed.barrier(group=slow_group)
    net = build_network(cfg, use_tensor_cores=use_tensor_cores).to(device)     # ONE instance shared by every block, like the reference
    ddp.broadcast_state(net)
    reducer = ddp.GradReducer(shard_optimizer=True) if world > 1 else None     # reduce-scatter, Adam on 1/world of the parameters, all-gather
    meters = (AverageMeter(), AverageMeter())
    if sh:
        n_frames = len(probe)
        tot_seg = int(np.ceil(n_frames / cfg.saveSegNum))
        scenes = vd.frame_size[cfg.dataset_name][-1]

        def seg(i, kind):
            return np.load(cfg.path('foreground_train_{}_seg_{}-{}.npy'.format(m, i, kind)), allow_pickle=True)
        model_set = [[[[] for _ in range(cfg.w_block)] for _ in range(cfg.h_block)] for _ in range(scenes)]
        raw_scores_set = [[[[] for _ in range(cfg.w_block)] for _ in range(cfg.h_block)] for _ in range(scenes)]
        of_scores_set = [[[[] for _ in range(cfg.w_block)] for _ in range(cfg.h_block)] for _ in range(scenes)]
        for s in range(scenes):
            for hh in range(cfg.h_block):
                for ww in range(cfg.w_block):
                    meters = (AverageMeter(), AverageMeter())

                    def batches(epoch, s=s, hh=hh, ww=ww):
                        for si in range(tot_seg):        # segments streamed from disk every epoch (train.py:293-299)
                            raw_seg = seg(si, 'raw')[s][hh][ww]
                            if len(raw_seg) == 0:        # a segment without cubes for this block: nothing to train on
                                continue
                            store = vd.DeviceCubeStore(raw_seg, seg(si, 'flow')[s][hh][ww], device=device)
                            yield from _epoch_batches(store, cfg, rank, world)
                    train_block(net, batches, cfg, reducer, meters, (s, hh, ww))
                    model_set[s][hh][ww].append(_state_dict_for_disk(net))
                    for si in range(tot_seg):
                        raw_seg = seg(si, 'raw')[s][hh][ww]
                        if len(raw_seg) == 0:            # the reference concatenates only what the segments hold
                            continue
                        store = vd.DeviceCubeStore(raw_seg, seg(si, 'flow')[s][hh][ww], device=device)
                        r, o = score_block(net, store, cfg.batch_size, cfg.useFlow)
                        raw_scores_set[s][hh][ww].append(r)
                        if cfg.useFlow:
                            of_scores_set[s][hh][ww].append(o)
                    if raw_scores_set[s][hh][ww]:
                        raw_scores_set[s][hh][ww] = np.concatenate(raw_scores_set[s][hh][ww], axis=0)
                        if cfg.useFlow:
                            of_scores_set[s][hh][ww] = np.concatenate(of_scores_set[s][hh][ww], axis=0)
    else:
        fs = np.load(cfg.path('foreground_train_{}-raw.npy'.format(m)), allow_pickle=True)
        fs2 = np.load(cfg.path('foreground_train_{}-flow.npy'.format(m)), allow_pickle=True)
        print('foreground for training data loaded!')
        model_set = [[[] for _ in range(len(fs[hh]))] for hh in range(len(fs))]
        raw_scores_set = [[[] for _ in range(len(fs[hh]))] for hh in range(len(fs))]
        of_scores_set = [[[] for _ in range(len(fs[hh]))] for hh in range(len(fs))]
        for hh in range(len(fs)):
            for ww in range(len(fs[hh])):
                if len(fs[hh][ww]) > 1:                   # "num > 1 for data parallel" (train.py:370)
                    store = vd.DeviceCubeStore(fs[hh][ww], fs2[hh][ww], device=device)     # the whole block on every rank
                    train_block(net, lambda epoch, store=store: _epoch_batches(store, cfg, rank, world), cfg, reducer, meters, (hh, ww))
                    model_set[hh][ww].append(_state_dict_for_disk(net))
                    raw_scores_set[hh][ww], of_scores_set[hh][ww] = score_block(net, store, cfg.batch_size, cfg.useFlow)
    if rank == 0:
        torch.save(raw_scores_set, cfg.path('raw_training_scores_{}.npy'.format(cfg.tag())))
        torch.save(of_scores_set, cfg.path('of_training_scores_{}.npy'.format(cfg.tag())))
        print('training scores saved!')
        torch.save(model_set, cfg.path('model_{}.npy'.format(cfg.tag())))
        print('Training of {} for dataset: {} has completed!'.format(cfg.method, cfg.dataset_name))
    return model_set


# ------------------------------------------------------------------------------------------ stage 4: scoring + evaluation
def _stats(scores):
    return (np.mean(scores), np.std(scores)) if len(scores) else (0.0, 1.0)


@torch.no_grad()
def score_frames(cfg, net_for, stats_for, foreground_set, foreground_set2, foreground_bbox_set, device, out_dir=None, scene_idx=None,
                 max_cubes=4096, score_batch=512):
    """test.py:270-358: per frame, per block: forward all cubes of the block, per-cube SSE, z-normalise with the training
    statistics, weight, paint the bbox rectangles with a running max.  Returns the list of per-frame score masks.

    The reference forwards one tiny batch per (frame, block) (about 17 cubes per frame on UCSDped2: launch-bound).  Eval-mode
    scores are per-cube and independent of what else is in the batch -- bit for bit, tests/test_unet_gpu.py::
    test_full_batch_properties -- so the cubes of consecutive frames that use the same model are scored in batches of up to
    ``max_cubes`` and scattered back; the masks are identical to frame-by-frame scoring.  Each flush runs through the engine in
    fixed sub-batches of ``score_batch`` cubes, so the activation workspace is bounded (28.5 MB per cube for 5raw5of: 14.6 GB at
    512) and never re-allocated for a slightly larger flush."""
    h, w = vd.frame_size[cfg.dataset_name][:2]
    n_frames = len(foreground_set)
    scores = [[[None for _ in foreground_set[f][hh]] for hh in range(len(foreground_set[f]))] for f in range(n_frames)]
    pending = {}          # model key -> list of (f, hh, ww, n)

    def flush(key):
        items = pending.pop(key, [])
        if not items:
            return
        s_, hh, ww = key
        net = net_for(s_, hh, ww)
        raw_np = np.concatenate([foreground_set[f][a][b] for (f, a, b, _) in items], axis=0)
        flow_np = np.concatenate([foreground_set2[f][a][b] for (f, a, b, _) in items], axis=0)
        raw_parts, of_parts = [], []
        for o in range(0, raw_np.shape[0], score_batch):
            x, x_of = vd.cubes_to_device_tensors(torch.as_tensor(raw_np[o:o + score_batch]).to(device),
                                                 torch.as_tensor(flow_np[o:o + score_batch]).to(device, torch.float32))
            r_, o_ = net.score(x, x_of)
            raw_parts.append(r_.cpu().numpy())
            if o_ is not None:
                of_parts.append(o_.cpu().numpy())
        (rm, rs), (om, os_) = stats_for(s_, hh, ww)
        sc = cfg.w_raw * ((np.concatenate(raw_parts) - rm) / rs)
        if cfg.useFlow:
            sc = sc + cfg.w_of * ((np.concatenate(of_parts) - om) / os_)
        o = 0
        for (f, a, b, n) in items:
            scores[f][a][b] = sc[o:o + n]
            o += n

    for f in range(n_frames):
        s_ = scene_idx[f] - 1 if scene_idx is not None else None
        for hh in range(len(foreground_set[f])):
            for ww in range(len(foreground_set[f][hh])):
                cubes = foreground_set[f][hh][ww]
                if len(cubes) == 0:
                    continue
                key = (s_, hh, ww)
                if net_for(s_, hh, ww) is None:          # objects where training saw none: anomaly (test.py:307-309)
                    scores[f][hh][ww] = np.ones(cubes.shape[0]) * BIG_NUMBER
                    continue
                pending.setdefault(key, []).append((f, hh, ww, cubes.shape[0]))
                if sum(it[3] for it in pending[key]) >= max_cubes:
                    flush(key)
    for key in list(pending):
        flush(key)
    masks = []
    for f in range(n_frames):
        if out_dir is not None:
            print('Calculating scores for {}-th frame'.format(f))
        pix = -1 * np.ones(shape=(h, w)) * BIG_NUMBER
        for hh in range(len(foreground_set[f])):
            for ww in range(len(foreground_set[f][hh])):
                if scores[f][hh][ww] is not None:
                    paint_score_mask(pix, scores[f][hh][ww], foreground_bbox_set[f][hh][ww], BIG_NUMBER)
        if out_dir is not None:
            torch.save(pix, os.path.join(out_dir, '{}'.format(f)))
        masks.append(pix)
    return masks


def test(cfg_path='config.cfg', results_dir='results', use_tensor_cores=None):
    if use_tensor_cores is None:
        use_tensor_cores = default_precision()
    cfg = Config(cfg_path, 'test')
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(device)
    probe = vd.unified_dataset_interface(cfg.dataset_name, os.path.join(cfg.raw_dataset_dir, cfg.dataset_name), context_frame_num=1,
                                         mode=cfg.mode, border_mode='hard')
    all_bboxes = load_or_make_bboxes(cfg, probe)
    m = cfg.foreground_extraction_mode
    sh = cfg.dataset_name == 'ShanghaiTech'
    if not cfg.foreground_saved:
        fs, fs2, fb, scene_idx = extract_foreground_test(cfg, all_bboxes)
    else:
        scene_idx = np.load(cfg.path('scene_idx.npy')) if sh else None
        fs = np.load(cfg.path('foreground_test_{}-raw.npy'.format(m)), allow_pickle=True)
        fs2 = np.load(cfg.path('foreground_test_{}-flow.npy'.format(m)), allow_pickle=True)
        fb = np.load(cfg.path('foreground_bbox_test_{}.npy'.format(m)), allow_pickle=True)
        print('foreground for testing data loaded!')
    mask_dir = os.path.join(results_dir, cfg.dataset_name, 'score_mask')
    if not cfg.scores_saved:
        os.makedirs(mask_dir, exist_ok=True)
        weights = torch.load(cfg.path('model_{}.npy'.format(cfg.tag())), weights_only=False)
        raw_tr = torch.load(cfg.path('raw_training_scores_{}.npy'.format(cfg.tag())), weights_only=False)
        of_tr = torch.load(cfg.path('of_training_scores_{}.npy'.format(cfg.tag())), weights_only=False)
        nets = {}
        pool = vu.WorkspacePool()              # ONE activation workspace for all per-block models (they are scored one at a time)

        def pick(tree, s, hh, ww):
            return tree[s][hh][ww] if sh else tree[hh][ww]

        def net_for(s, hh, ww):
            key = (s, hh, ww)
            if key not in nets:
                sd = pick(weights, s, hh, ww)
                if len(sd) == 0:
                    nets[key] = None
                else:
                    net = build_network(cfg, use_tensor_cores=use_tensor_cores)
                    load_block_state(net, sd[0])
                    nets[key] = net.to(device).eval().share_workspace(pool)
            return nets[key]

        def stats_for(s, hh, ww):
            return _stats(pick(raw_tr, s, hh, ww)), (_stats(pick(of_tr, s, hh, ww)) if cfg.useFlow else (0.0, 1.0))
        score_frames(cfg, net_for, stats_for, fs, fs2, fb, device, out_dir=mask_dir, scene_idx=scene_idx)
    return evaluate(cfg, results_dir, scene_idx)


def evaluate(cfg, results_dir='results', scene_idx=None):
    """Frame-level AUROC from the saved score masks (test.py:362-399)."""
    from torch.utils.data import DataLoader
    ds = vd.unified_dataset_interface(cfg.dataset_name, os.path.join(cfg.raw_dataset_dir, cfg.dataset_name), context_frame_num=0,
                                      mode=cfg.mode, border_mode='hard')
    loader = DataLoader(dataset=ds, batch_size=1, shuffle=False, num_workers=0, collate_fn=vd.bbox_collate(cfg.mode).collate)
    print('Evaluating {} by frame-criterion:'.format(cfg.dataset_name))
    sh = cfg.dataset_name == 'ShanghaiTech'
    n_scene = ds.scene_num if sh else 1
    frame_scores, targets = [[] for _ in range(n_scene)], [[] for _ in range(n_scene)]
    for idx, (_, target) in enumerate(loader):
        pix = torch.load(os.path.join(results_dir, cfg.dataset_name, 'score_mask', '{}'.format(idx)), weights_only=False)
        s = scene_idx[idx] - 1 if sh else 0
        frame_scores[s].append(pix.max())
        targets[s].append(target[0].numpy().max())
    aucs = []
    for s in range(n_scene):
        name = '{}_{}_{}_frame_results{}.npz'.format(cfg.modality, cfg.foreground_extraction_mode, cfg.method, '_scene_%d' % (s + 1) if sh else '')
        path = os.path.join(results_dir, cfg.dataset_name, name)
        print('Results written to {}:'.format(path))
        aucs.append(save_roc_pr_curve_data(np.array(frame_scores[s]), np.array(targets[s]) > 0, path))
    if sh:
        print('Average frame-level AUC is {}'.format(np.array(aucs).mean()))
    return float(np.array(aucs).mean())
