"""Optical-flow pre-computation stage (reference: calc_optical_flow.py:12-88) on vec_vad_b200.flownet2.FlowNet2.

For every frame of a dataset (``unified_dataset_interface(..., context_frame_num=1, border_mode='hard')``, i.e. the window
[previous, current, next] clamped inside the video) the flow of one frame pair is computed at 512x384 and resized back to the
frame size: the pair is (window[0], window[1]) at a video border -- where the clamped window repeats a frame, so the very first
frame of a video is paired with itself -- and (window[1], window[2]) everywhere else (calc_optical_flow.py:44,61).  The result is saved as ``<name>.npy`` under ``./optical_flow``
in the directory layout of ``./raw_datasets``; ``vad_datasets`` reads those files back as the flow modality.

Host glue (cv2 resize, file naming) follows the reference; the network runs on the CUDA path only.
"""
import os

import numpy as np
import torch

from . import flownet2 as fn
from . import vad_datasets as vd

NET_SIZE = (512, 384)                      # (width, height) handed to cv2.resize (calc_optical_flow.py:49-54)


def load_flownet2(checkpoint='FlowNet2_src/pretrained/FlowNet2_checkpoint.pth.tar', device='cuda'):
    """FlowNet2 with the published checkpoint loaded the way calc_optical_flow.py:15-22 does (keys filtered to the model's own)."""
    net = fn.FlowNet2()
    pretrained = torch.load(checkpoint, map_location='cpu')['state_dict']
    sd = net.state_dict()
    sd.update({k: v for k, v in pretrained.items() if k in sd})
    net.load_state_dict(sd)
    return net.to(device).eval()


def frame_pair(cur_imgs, frame_range):
    """cur_imgs [3,H,W,C] (the 'hard' window) -> the two frames fed to the network, resized to 512x384, 3 channels, float32
    [1,3,2,384,512]                                                             (calc_optical_flow.py:44-56,61-74)"""
    import cv2
    border = frame_range[1] == frame_range[0] or frame_range[1] == frame_range[2]
    a, b = (cur_imgs[0], cur_imgs[1]) if border else (cur_imgs[1], cur_imgs[2])
    if cur_imgs.shape[3] == 1:
        im1 = np.concatenate([cv2.resize(a, NET_SIZE)[:, :, np.newaxis]] * 3, axis=2)
        im2 = np.concatenate([cv2.resize(b, NET_SIZE)[:, :, np.newaxis]] * 3, axis=2)
    else:
        im1, im2 = cv2.resize(a, NET_SIZE), cv2.resize(b, NET_SIZE)
    return np.array([[im1, im2]]).transpose((0, 4, 1, 2, 3)).astype(np.float32)


def calc_optical_flow(dataset, net=None, of_root_dir='./optical_flow', checkpoint='FlowNet2_src/pretrained/FlowNet2_checkpoint.pth.tar',
                      verbose=True, batch_pairs=1):
    """calc_optical_flow.py:12-88.  ``net``: a FlowNet2 already on the GPU (default: load the published checkpoint).

    ``batch_pairs`` > 1 sends that many frame pairs through the network in one call (the reference feeds one pair at a time, which is the
    default here): the deep layers' small pixel grids then fill the GPU (6.0 instead of 7.8 ms per pair at 8).  No
    kernel couples the images of a batch, but the conv planner picks tile shapes and splits of the contraction by the batched shape,
    so a batched flow can differ from the one-pair flow in the last bits (summation order); files and naming are the same."""
    import cv2
    if net is None:
        net = load_flownet2(checkpoint)
    dev = next(net.parameters()).device
    depth = len(dataset.dir.split('/')) - 1
    pending = []                                             # (output file, frame size, network input [1,3,2,384,512])

    def flush():
        if not pending:
            return
        ims = torch.from_numpy(np.concatenate([p[2] for p in pending])).to(dev)
        pred = net(ims).cpu().numpy()
        for (path, old_size, _), flow in zip(pending, pred):
            np.save(path, cv2.resize(flow.transpose((1, 2, 0)), old_size))
        del pending[:]

    for idx in range(len(dataset)):
        if verbose:
            print('Calculating optical flow for {}-th frame'.format(idx + 1))
        addr = dataset.all_frame_addr[idx]
        name = addr.split('/')[-1].split('.')[0]
        of_path = os.path.join(of_root_dir, *addr.split('/')[depth:-1])
        os.makedirs(of_path, exist_ok=True)
        batch = dataset[idx][0]
        cur_imgs = np.transpose(batch.cpu().numpy(), [0, 2, 3, 1])
        old_size = (cur_imgs.shape[2], cur_imgs.shape[1])
        pending.append((os.path.join(of_path, name + '.npy'), old_size, frame_pair(cur_imgs, dataset.context_range(idx))))
        if len(pending) >= max(1, int(batch_pairs)):
            flush()
    flush()


def main(dataset_name='UCSDped2', net=None, batch_pairs=None):
    """The reference script body (calc_optical_flow.py:108-114): training set, then testing set.  ``batch_pairs`` (default: the
    environment's VECVAD_FLOW_BATCH, else 1 = the reference's one pair per call) as in calc_optical_flow."""
    if batch_pairs is None:
        batch_pairs = int(os.environ.get('VECVAD_FLOW_BATCH', '1'))
    for mode in ('train', 'test'):
        ds = vd.unified_dataset_interface(dataset_name=dataset_name, dir=os.path.join('raw_datasets', dataset_name), context_frame_num=1,
                                          mode=mode, border_mode='hard')
        calc_optical_flow(ds, net=net, batch_pairs=batch_pairs)
