"""Writes tests/golden/crop_resize.npz: seeded frame stacks + boxes and the patches the REFERENCE's get_foreground
(/root/reference/vad_datasets.py:70-93, imported unmodified; it calls cv2.resize) cuts from them.  Run in the build container:
    python tests/golden/make_crop_resize_golden.py
"""
import os
import sys

import numpy as np

np.int = int                       # numpy >= 1.24 dropped the alias the reference still uses (vad_datasets.py:74-75)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('VECVAD_REFERENCE', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import vad_datasets as ref         # noqa: E402  (the reference module)


def boxes_for(rng, H, W, n):
    """Float boxes like bboxes_*.npy: fractional edges, sizes from 1 px to most of the frame; fixed special sizes first."""
    out = []
    special = [(32, 32), (64, 64), (1, 1), (2, 5), (31, 33), (64, 32), (96, 96), (17, 90), (128, 128)]
    for (h, w) in special:
        if h < H and w < W:
            y0, x0 = int(rng.integers(0, H - h)), int(rng.integers(0, W - w))
            out.append([x0 - 0.3, y0 - 0.6, x0 + w - 0.2, y0 + h - 0.5])      # ceil -> exactly (w, h)
    while len(out) < n:
        x0, y0 = rng.uniform(0, W - 2), rng.uniform(0, H - 2)
        x1, y1 = rng.uniform(x0 + 1, W - 0.01), rng.uniform(y0 + 1, H - 0.01)
        out.append([x0, y0, x1, y1])
    return np.array(out, dtype=np.float32)


def main():
    rng = np.random.default_rng(20261017)
    d = {}
    H, W = 150, 200                                        # small frames keep the fixture under 1 MB
    raw = rng.integers(0, 256, (2, 3, H, W), dtype=np.uint8)
    flow = (rng.standard_normal((1, 2, H, W)) * 3).astype(np.float32)
    boxes = boxes_for(rng, H, W, 32)
    d['raw'], d['flow'], d['boxes'] = raw, flow, boxes
    d['raw_patches'] = ref.get_foreground(raw, boxes, 32)
    d['flow_patches'] = ref.get_foreground(flow, boxes, 32)
    d['raw3_patches'] = ref.get_foreground(raw[0], boxes, 32)          # the 3-D branch (context_frame_num = 0)
    d['raw_patches_p16'] = ref.get_foreground(raw, boxes[:10], 16)
    np.savez_compressed(os.path.join(HERE, 'crop_resize.npz'), **d)
    print({k: (v.shape, str(v.dtype)) for k, v in d.items()})


if __name__ == '__main__':
    main()
