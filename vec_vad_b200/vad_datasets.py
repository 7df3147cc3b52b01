"""Dataset surface of the reference's ``vad_datasets.py`` for the hot path.

Kept API (same names, argument meaning, return types):
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
import ctypes as C

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
    return np.transpose(img_batch, axes).numpy()


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


# ------------------------------------------------------------------------------------------ device feed
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
