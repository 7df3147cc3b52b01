"""Dataset surface of the reference's ``vad_datasets.py`` for the hot path.

Kept API (same names, argument meaning, return types):
  * ``unified_dataset_interface`` + ``ped_dataset`` / ``avenue_dataset`` / ``shanghaiTech_dataset``     vad_datasets.py:95-114,170,404,613
    (frame-folder datasets; host I/O, not accelerated -- their integer index paths are restated bit-exactly:
    ``context_range`` vad_datasets.py:277-354, bbox ceil/crop :74-75)
  * ``bbox_collate``, ``get_inputs``                                                                    vad_datasets.py:18-25,48-68
  * ``cube_to_train_dataset(data, target=None, tranform=transform)``   vad_datasets.py:130-168
    (the misspelt ``tranform`` keyword is the reference's)
  * ``transform`` / ``ToTensor``                                        vad_datasets.py:12-14
  * ``get_foreground(img, bboxes, patch_size)``                         vad_datasets.py:70-93
  * ``img_tensor2numpy`` / ``img_batch_tensor2numpy``                   vad_datasets.py:27-46
  * ``frame_size``                                                      vad_datasets.py:16
New, device side (the feed for >100k STC/s):
  * ``cubes_to_device_tensors(raw_u8, flow)`` -- uint8 cubes already in HBM -> the float tensors
    ``cube_to_train_dataset`` + ``DataLoader`` collate would have produced, in one kernel.
  * ``DeviceCubeStore`` -- all cubes of a block resident in HBM as uint8 (15 KB/STC) + fp32 flow,
    shuffled mini-batches gathered and converted on the device.
"""
import glob
import os
from collections import OrderedDict

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib

# (h, w, file_format, scene_num)                                        vad_datasets.py:16
frame_size = {'UCSDped1': (158, 238, '.tif', 1), 'UCSDped2': (240, 360, '.tif', 1), 'avenue': (360, 640, '.jpg', 1),
              'ShanghaiTech': (480, 856, '.jpg', 1)}


class ToTensor:
    """torchvision.transforms.ToTensor for ndarrays: HWC -> CHW; uint8 is converted to float32 and divided by 255,
    every other dtype is passed through unchanged."""

    def __call__(self, pic):
        if not isinstance(pic, np.ndarray):
            raise TypeError('pic should be ndarray, got %s' % type(pic))
        if pic.ndim == 2:
            pic = pic[:, :, None]
        t = torch.from_numpy(np.ascontiguousarray(pic.transpose((2, 0, 1))))
        if t.dtype == torch.uint8:
            return t.to(torch.float32).div(255)
        return t


transform = ToTensor()


def img_tensor2numpy(img):
    if isinstance(img, np.ndarray):
        return torch.from_numpy(np.transpose(img, [2, 0, 1]))
    return np.transpose(img, [1, 2, 0]).numpy()


def img_batch_tensor2numpy(img_batch):
    if isinstance(img_batch, np.ndarray):
        axes = [0, 3, 1, 2] if img_batch.ndim == 4 else [0, 1, 4, 2, 3]
        return torch.from_numpy(np.transpose(img_batch, axes))
    axes = [0, 2, 3, 1] if img_batch.dim() == 4 else [0, 1, 3, 4, 2]
    return np.transpose(img_batch.cpu(), axes).numpy()


def _bbox_to_slices(box):
    """ceil on all four edges, then int (vad_datasets.py:74-75): the integer path that must stay bit-exact."""
    x_min, x_max = int(np.ceil(box[0])), int(np.ceil(box[2]))
    y_min, y_max = int(np.ceil(box[1])), int(np.ceil(box[3]))
    return slice(y_min, y_max), slice(x_min, x_max)


def get_foreground(img, bboxes, patch_size):
    """Crop every bbox from a [C,H,W] frame or a [T,C,H,W] frame stack and resize to patch_size (cv2 INTER_LINEAR)."""
    import cv2
    patches = []
    if img.ndim == 3:
        for box in bboxes:
            ys, xs = _bbox_to_slices(box)
            p = cv2.resize(np.transpose(img[:, ys, xs], [1, 2, 0]), (patch_size, patch_size))
            patches.append(np.transpose(p, [2, 0, 1]))
    elif img.ndim == 4:
        for box in bboxes:
            ys, xs = _bbox_to_slices(box)
            cube = [np.transpose(cv2.resize(np.transpose(img[j][:, ys, xs], [1, 2, 0]), (patch_size, patch_size)), [2, 0, 1])
                    for j in range(img.shape[0])]
            patches.append(np.array(cube))
    return np.array(patches)


def _fold_time(cube):
    """[T,H,W,C] -> [H,W,T*C]  (transpose [1,2,0,3] + reshape, vad_datasets.py:159-160)"""
    c = np.transpose(cube, [1, 2, 0, 3])
    return np.reshape(c, (c.shape[0], c.shape[1], -1))


class cube_to_train_dataset(Dataset):
    def __init__(self, data, target=None, tranform=transform):
        if data.ndim == 4:
            data = data[:, np.newaxis]
        if target is not None and target.ndim == 4:    # (the reference dereferences target.shape even when it is None)
            target = target[:, np.newaxis]
        self.data, self.target, self.transform = data, target, tranform

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, indice):
        tf = self.transform if self.transform is not None else (lambda a: a)
        cur = self.data[indice]
        if self.target is None:
            return tf(_fold_time(cur[:-1])), tf(cur[-1])
        return tf(_fold_time(cur)), tf(_fold_time(self.target[indice])), tf(_fold_time(cur.copy()))


# ------------------------------------------------------------------------------------------ frame-folder datasets
def get_inputs(file_addr):
    ext = file_addr.split('.')[-1]
    if ext == 'mat':
        import scipy.io as sio
        return sio.loadmat(file_addr, verify_compressed_data_integrity=False)['uv']
    if ext == 'npy':
        return np.load(file_addr)
    import cv2
    return cv2.imread(file_addr)


def context_range(indice, frame_video_idx, context_frame_num, border_mode):
    """Indices of the frames forming the temporal context of frame ``indice`` (vad_datasets.py:277-354; integer path, bit-exact).

    ``frame_video_idx[i]`` is the 1-based video a frame belongs to.  A context never crosses a video boundary: missing frames
    are replaced by repeating the first / last frame of the same video ('predict' and 'hard'), or the window slides inside
    the video ('elastic').  Raises NotImplementedError (like the reference) when a video is shorter than the context."""
    n, c = len(frame_video_idx), context_frame_num
    if border_mode == 'elastic':
        if indice - c < 0:
            indice = c
        elif indice + c > n - 1:
            indice = n - 1 - c
        start, end, need = indice - c, indice + c, 2 * c + 1
    elif border_mode == 'predict':
        start, end, need = max(indice - c, 0), indice, c + 1
    else:
        start, end, need = max(indice - c, 0), min(indice + c, n - 1), 2 * c + 1
    center = frame_video_idx[indice]
    vids = list(frame_video_idx[start:end + 1])
    pad = need - len(vids)
    if pad > 0:
        vids = [vids[0]] * pad + vids if start == 0 else vids + [vids[-1]] * pad
    rel = np.array(vids) - center
    offset = int(rel.sum())
    if rel[0] != 0 and rel[-1] != 0:
        raise NotImplementedError('The video is too short or the context frame number is too large!')
    if pad == 0 and offset == 0:
        return list(range(start, end + 1))
    if border_mode == 'elastic':
        return list(range(start - offset, end - offset + 1))
    if pad > 0 and abs(offset) > 0:
        raise NotImplementedError('The video is too short or the context frame number is too large!')
    if border_mode == 'predict':
        idx = list(range(start - offset, end + 1))
        return [idx[0]] * max(abs(offset), pad) + idx
    if offset > 0:
        idx = list(range(start, end - offset + 1))
        return idx + [idx[-1]] * offset
    if offset < 0:
        idx = list(range(start - offset, end + 1))
        return [idx[0]] * (-offset) + idx
    if start == 0:
        idx = list(range(start, end + 1))
        return [idx[0]] * pad + idx
    idx = list(range(start, end + 1))
    return idx + [idx[-1]] * pad


class bbox_collate:
    def __init__(self, mode):
        self.mode = mode

    def collate(self, batch):
        data, target = [x[0] for x in batch], [x[1] for x in batch]
        if self.mode == 'train':
            return torch.cat(data, dim=0), target
        if self.mode == 'test':
            return data, target
        raise NotImplementedError


class _frame_folder_dataset(Dataset):
    """Shared body of the three reference dataset classes (which are copies of each other apart from the directory
    layout and the ground-truth format).  ``__getitem__`` -> (img_batch, gt-or-zeros(1))."""

    def __init__(self, dir, mode='train', context_frame_num=0, border_mode='elastic', file_format=None, all_bboxes=None, patch_size=32):
        self.dir, self.mode = dir, mode
        self.videos = OrderedDict()
        self.all_frame_addr, self.frame_video_idx = [], []
        self.tot_frame_num = 0
        self.context_frame_num, self.border_mode = context_frame_num, border_mode
        self.file_format, self.all_bboxes, self.patch_size = file_format, all_bboxes, patch_size
        self.return_gt = False
        self.foreground_device = None          # a CUDA device: __getitem__ crops + resizes the boxes on the GPU (get_foreground_device)
        if mode not in ('train', 'test'):
            raise NotImplementedError
        self.dataset_init()

    def __len__(self):
        return self.tot_frame_num

    def _add_videos(self, video_dirs, name_filter=None, per_video=None):
        idx = 1
        for video in sorted(video_dirs):
            name = video.split('/')[-1]
            if name_filter is not None and name_filter not in name:
                continue
            frames = sorted(glob.glob(os.path.join(video, '*' + self.file_format)))
            self.videos[name] = {'path': video, 'frame': frames, 'length': len(frames)}
            self.frame_video_idx += [idx] * len(frames)
            idx += 1
            if per_video is not None:
                per_video(name, len(frames))
        for cont in self.videos.values():
            self.all_frame_addr += cont['frame']
        self.tot_frame_num = len(self.all_frame_addr)

    def context_range(self, indice):
        return context_range(indice, self.frame_video_idx, self.context_frame_num, self.border_mode)

    def _gt(self, indice):
        raise NotImplementedError

    def __getitem__(self, indice):
        if self.context_frame_num == 0:
            img = np.transpose(get_inputs(self.all_frame_addr[indice]), [2, 0, 1])
        else:
            img = np.array([np.transpose(get_inputs(self.all_frame_addr[i]), [2, 0, 1]) for i in self.context_range(indice)])
        if self.all_bboxes is not None and self.foreground_device is not None:
            # device path (f4): upload the frame stack once, crop + resize every box on the GPU (bit-identical patches)
            img = get_foreground_device(torch.from_numpy(np.ascontiguousarray(img)).to(self.foreground_device), self.all_bboxes[indice],
                                        self.patch_size)
        else:
            if self.all_bboxes is not None:
                img = get_foreground(img=img, bboxes=self.all_bboxes[indice], patch_size=self.patch_size)
            img = torch.from_numpy(img)
        if self.mode == 'test' and self.return_gt:
            return img, torch.from_numpy(self._gt(indice))
        return img, torch.zeros(1)


class ped_dataset(_frame_folder_dataset):                      # UCSD ped1 / ped2          vad_datasets.py:170-402
    def __init__(self, dir, mode='train', context_frame_num=0, border_mode='elastic', file_format='.tif', all_bboxes=None, patch_size=32):
        self.h, self.w = (158, 238) if dir[-1] == '1' else (240, 360)
        self.all_gt_addr, self.gts = [], OrderedDict()
        super().__init__(dir, mode, context_frame_num, border_mode, file_format, all_bboxes, patch_size)

    def dataset_init(self):
        if self.mode == 'train':
            self._add_videos(glob.glob(os.path.join(self.dir, 'Train', '*')), name_filter='Train')
            return
        entries = sorted(glob.glob(os.path.join(self.dir, 'Test', '*')))
        gt_dirs = [d for d in entries if '_gt' in d]
        self.return_gt = len(gt_dirs) > 0
        self._add_videos([d for d in entries if '_gt' not in d and 'Test' in d.split('/')[-1]])
        for gt in gt_dirs:
            frames = sorted(glob.glob(os.path.join(gt, '*.bmp')))
            self.gts[gt.split('/')[-1]] = {'gt_frame': frames}
            self.all_gt_addr += frames

    def _gt(self, indice):
        import cv2
        return cv2.imread(self.all_gt_addr[indice], cv2.IMREAD_GRAYSCALE)


class avenue_dataset(_frame_folder_dataset):                   # vad_datasets.py:404-611
    def __init__(self, dir, mode='train', context_frame_num=0, border_mode='elastic', file_format='.jpg', all_bboxes=None, patch_size=32):
        self.all_gt = []
        super().__init__(dir, mode, context_frame_num, border_mode, file_format, all_bboxes, patch_size)

    def dataset_init(self):
        sub = 'training' if self.mode == 'train' else 'testing'
        self._add_videos(glob.glob(os.path.join(self.dir, sub, 'frames', '*')))
        if self.mode == 'test':
            gt_dir = os.path.join(self.dir, 'ground_truth_demo', 'testing_label_mask')
            if os.path.exists(gt_dir):
                import scipy.io as sio
                self.return_gt = True
                vols = [sio.loadmat(os.path.join(gt_dir, str(x + 1) + '_label.mat'))['volLabel'] for x in range(len(self.videos))]
                self.all_gt = np.concatenate(vols, axis=1)

    def _gt(self, indice):
        return self.all_gt[0, indice]


class shanghaiTech_dataset(_frame_folder_dataset):             # vad_datasets.py:613-835
    def __init__(self, dir, mode='train', context_frame_num=0, border_mode='elastic', file_format='.jpg', all_bboxes=None, patch_size=32):
        self.save_scene_idx, self.scene_idx, self.scene_num, self.all_gt = [], [], 0, []
        super().__init__(dir, mode, context_frame_num, border_mode, file_format, all_bboxes, patch_size)

    def dataset_init(self):
        def per_video(name, nframes):
            self.save_scene_idx += [int(name[:2])] * nframes      # frames are saved by scene
            self.scene_idx += [1] * nframes                       # ... and processed as one scene
        if self.mode == 'train':
            dirs = sorted(glob.glob(os.path.join(self.dir, 'training', 'videosFrame', '*')))
        else:   # the two test parts are each sorted, part 1 first (vad_datasets.py:679-681)
            base = os.path.join(self.dir, 'Testing', 'frames_part')
            dirs = sorted(glob.glob(os.path.join(base + '1', '*'))) + sorted(glob.glob(os.path.join(base + '2', '*')))
        self._add_videos_in_order(dirs, per_video)
        self.scene_num = len(set(self.scene_idx))
        if self.mode == 'test':
            gt_dir = os.path.join(self.dir, 'Testing', 'test_frame_mask')
            if os.path.exists(gt_dir):
                self.return_gt = True
                self.all_gt = np.concatenate([np.load(g) for g in sorted(glob.glob(os.path.join(gt_dir, '*')))], axis=0)

    def _add_videos_in_order(self, dirs, per_video):
        idx = 1
        for video in dirs:
            name = video.split('/')[-1]
            frames = sorted(glob.glob(os.path.join(video, '*' + self.file_format)))
            self.videos[name] = {'path': video, 'frame': frames, 'length': len(frames)}
            self.frame_video_idx += [idx] * len(frames)
            idx += 1
            per_video(name, len(frames))
        for cont in self.videos.values():
            self.all_frame_addr += cont['frame']
        self.tot_frame_num = len(self.all_frame_addr)

    def _gt(self, indice):
        return np.array([self.all_gt[indice]])


def unified_dataset_interface(dataset_name, dir, mode='train', context_frame_num=0, border_mode='elastic', file_format=None,
                              all_bboxes=None, patch_size=32):
    if file_format is None:
        if dataset_name in ('UCSDped1', 'UCSDped2'):
            file_format = '.tif'
        elif dataset_name in ('avenue', 'ShanghaiTech'):
            file_format = '.jpg'
        else:
            raise NotImplementedError
    cls = {'UCSDped1': ped_dataset, 'UCSDped2': ped_dataset, 'avenue': avenue_dataset, 'ShanghaiTech': shanghaiTech_dataset}.get(dataset_name)
    if cls is None:
        raise NotImplementedError
    return cls(dir=dir, context_frame_num=context_frame_num, mode=mode, border_mode=border_mode, all_bboxes=all_bboxes,
               patch_size=patch_size, file_format=file_format)


# ------------------------------------------------------------------------------------------ device feed
def get_foreground_device(img, bboxes, patch_size, layout='CHW'):
    """``get_foreground`` (vad_datasets.py:70-93) on the GPU: same boxes -> the same patches, bit for bit (uint8 AND float32).

    img    : CUDA tensor, uint8 or float32.  layout 'CHW': [C,H,W] or [T,C,H,W] as the reference passes it; layout 'HWC': [H,W,C]
             or [T,H,W,C] as cv2.imread / np.load deliver frames (no host transpose needed).
    bboxes : host array [N,4] (x_min, y_min, x_max, y_max) as loaded from bboxes_*.npy; the ceil -> int of every edge is done on
             the host in numpy exactly as the reference does it (the integer path stays the reference's own arithmetic).
    -> CUDA tensor [N,C,ps,ps] (3-D input) or [N,T,C,ps,ps] (4-D input), dtype of img.
    """
    _lib.require_cuda(img)
    if img.dtype not in (torch.uint8, torch.float32):
        raise TypeError('get_foreground_device: frames must be uint8 or float32, got %s' % img.dtype)
    if layout not in ('CHW', 'HWC'):
        raise ValueError("layout must be 'CHW' or 'HWC'")
    single = img.dim() == 3
    if img.dim() not in (3, 4):
        raise ValueError('img must be 3-D or 4-D, got %s' % (tuple(img.shape),))
    v = img[None] if single else img
    if layout == 'HWC':
        v = v.permute(0, 3, 1, 2)                            # a view: the kernel walks the frames through element strides
    T, Cn, H, W = v.shape
    boxes = np.asarray(bboxes)
    n = len(boxes)
    ib = np.empty((n, 4), dtype=np.int32)
    for k in range(n):
        ys, xs = _bbox_to_slices(boxes[k])
        # numpy slicing clips silently; the reference then hands cv2 whatever is left (an empty crop raises there)
        y0, y1, _ = ys.indices(H)
        x0, x1, _ = xs.indices(W)
        if y1 <= y0 or x1 <= x0:
            raise ValueError('get_foreground_device: box %d (%s) is empty after ceil / clipping to the %dx%d frame' % (k, boxes[k], H, W))
        ib[k] = (x0, y0, x1, y1)
    out = torch.empty((n, T, Cn, patch_size, patch_size), dtype=img.dtype, device=img.device)
    if n:
        dev_boxes = torch.from_numpy(ib).to(img.device, non_blocking=False)
        st = v.stride()
        _lib.check(_lib.lib().vecvad_crop_resize(_lib.ptr(v), int(img.dtype == torch.float32), T, Cn, H, W, st[0], st[1], st[2], st[3],
                                                 _lib.ptr(dev_boxes), n, int(patch_size), _lib.ptr(out), _lib.cur_stream()), 'crop_resize')
    return out[:, 0] if single else out


def cubes_to_device_tensors(raw_u8, flow=None):
    """uint8 cubes [N,T,S,S,3] (+ float32 flow [N,T_of,S,S,2]) on the GPU -> (x [N,3T,S,S] float32 in [0,1], x_of [N,2T_of,S,S]).

    Bit-identical to stacking ``cube_to_train_dataset`` items (uint8 -> float32 -> /255)."""
    _lib.require_cuda(raw_u8, flow)
    if raw_u8.dtype != torch.uint8 or raw_u8.dim() != 5 or raw_u8.shape[-1] != 3 or raw_u8.shape[2] != raw_u8.shape[3]:
        raise ValueError('raw cubes must be uint8 [N,T,S,S,3], got %s %s' % (raw_u8.dtype, tuple(raw_u8.shape)))
    raw_u8 = raw_u8.contiguous()
    n, t, s = raw_u8.shape[0], raw_u8.shape[1], raw_u8.shape[2]
    x = torch.empty((n, 3 * t, s, s), dtype=torch.float32, device=raw_u8.device)
    x_of, t_of = None, 0
    if flow is not None:
        if flow.dim() == 4:
            flow = flow[:, None]
        if flow.dtype != torch.float32 or flow.shape[0] != n or flow.shape[-1] != 2 or flow.shape[2] != s:
            raise ValueError('flow cubes must be float32 [N,T_of,S,S,2], got %s %s' % (flow.dtype, tuple(flow.shape)))
        flow = flow.contiguous()
        t_of = flow.shape[1]
        x_of = torch.empty((n, 2 * t_of, s, s), dtype=torch.float32, device=raw_u8.device)
    _lib.check(_lib.lib().vecvad_cubes_to_tensors(_lib.ptr(raw_u8), _lib.ptr(flow), _lib.ptr(x), _lib.ptr(x_of), n, t, t_of, s,
                                                  _lib.cur_stream()), 'cubes_to_tensors')
    return x, x_of


class DeviceCubeStore:
    """All STCs of one block resident in HBM (uint8 raw + fp32 flow); yields shuffled float mini-batches.

    Replaces ``DataLoader(cube_to_train_dataset(...), batch_size, shuffle=True)`` (train.py:373) for the fast path: the
    permutation comes from ``torch.randperm`` on the host generator exactly like ``RandomSampler``, so with the same seed
    the batches hold the same cubes in the same order as the reference loader."""

    def __init__(self, raw_u8, flow, device='cuda'):
        raw = torch.as_tensor(raw_u8)
        if raw.dim() == 4:
            raw = raw[:, None]
        fl = torch.as_tensor(flow)
        if fl.dim() == 4:
            fl = fl[:, None]
        self.raw = raw.to(device).contiguous()
        self.flow = fl.to(device, torch.float32).contiguous()

    def __len__(self):
        return self.raw.shape[0]

    def batches(self, batch_size, shuffle=True, generator=None, drop_last=False):
        n = len(self)
        if shuffle:
            if generator is None:
                seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler's own seeding
                generator = torch.Generator().manual_seed(seed)
            order = torch.randperm(n, generator=generator)
        else:
            order = torch.arange(n)
        order = order.to(self.raw.device)
        for i in range(0, n, batch_size):
            idx = order[i:i + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            yield cubes_to_device_tensors(self.raw.index_select(0, idx), self.flow.index_select(0, idx))

    def rank_batches(self, batch_size, rank, world, seed=None, shuffle=True):
        """Multi-process feed: yields ``(x, x_of, local_n, global_n)`` for EVERY global batch of the epoch (``ddp.rank_batch_plan``:
        one shared permutation, each global batch split over the ranks like DataParallel's scatter).  ``x`` is None when this
        rank's share of a ragged last batch is empty -- the caller must still take part in that step's gradient reduction.
        The store holds the whole block on every rank (15 KB + 8..40 KB per STC)."""
        from . import ddp
        for idx, global_n in ddp.rank_batch_plan(len(self), batch_size, rank, world, seed=seed, shuffle=shuffle):
            if idx.numel() == 0:
                yield None, None, 0, global_n
                continue
            idx = idx.to(self.raw.device)
            x, x_of = cubes_to_device_tensors(self.raw.index_select(0, idx), self.flow.index_select(0, idx))
            yield x, x_of, int(idx.numel()), global_n


class HostCubeFeeder:
    """Mini-batches that live in HOST memory (pinned uint8 cubes + fp32 flow) -> float device tensors, one batch ahead.

    For cube sets that are not kept resident in HBM (``DeviceCubeStore``), e.g. a block streamed from disk segment by segment
    (train.py:293-299): ``next()`` hands out batch i -- whose host-to-device copy was queued on a copy stream while batch i-1 was
    being trained on -- and immediately queues the copy of batch i+1, so the copy overlaps the step instead of preceding it.  The
    uint8 -> float conversion (``cube_to_train_dataset`` + ``ToTensor``) runs on the consumer's stream as in ``cubes_to_device_tensors``.

    batches: a sequence of ``(raw_u8 [B,T,S,S,3] uint8, flow [B,T_of,S,S,2] float32)`` host tensors, pinned for asynchronous copies;
    it is cycled (``next()`` never ends; the caller decides how many steps an epoch has)."""

    def __init__(self, batches, device='cuda'):
        self.batches = list(batches)
        if not self.batches:
            raise ValueError('HostCubeFeeder needs at least one batch')
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._i = 0
        self._pending = None
        self._queue()

    def _queue(self):
        raw, flow = self.batches[self._i % len(self.batches)]
        self._i += 1
        with torch.cuda.stream(self.copy_stream):
            d_raw = raw.to(self.device, non_blocking=True)
            d_flow = flow.to(self.device, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._pending = (d_raw, d_flow, done)

    def next(self):
        d_raw, d_flow, done = self._pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(done)
        d_raw.record_stream(cur)                 # allocated on the copy stream, consumed on the caller's
        d_flow.record_stream(cur)
        self._queue()                            # the next batch's copy overlaps the step the caller is about to run
        return cubes_to_device_tensors(d_raw, d_flow)
