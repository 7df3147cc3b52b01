"""A tiny UCSDped2-shaped dataset on disk (frames, FlowNet2-style flow .npy, ground truth, bboxes, config.cfg) for the
pipeline tests: moving bright squares on a dark background; 'anomalous' test frames contain a much faster square."""
import os

import numpy as np

H, W = 240, 360


def _frame(t, speed, seed):
    rng = np.random.RandomState(seed)
    img = (rng.rand(H, W) * 20).astype(np.uint8)
    boxes, flows = [], np.zeros((H, W, 2), np.float32)
    for k in range(3):
        x = int(20 + 100 * k + speed * t) % (W - 50)
        y = 40 + 50 * k
        img[y:y + 40, x:x + 30] = 150 + 30 * k
        boxes.append([x - 2.5, y - 2.5, x + 32.5, y + 42.5])
        flows[y:y + 40, x:x + 30, 0] = speed
    return np.stack([img] * 3, -1), np.array(boxes, np.float32), flows


def make(root, n_train=(12, 10), n_test=(10,), cfg_overrides=None):
    import cv2
    os.makedirs(root, exist_ok=True)
    tr_boxes, te_boxes = [], []
    for split, lens, bl in (('Train', n_train, tr_boxes), ('Test', n_test, te_boxes)):
        for v, n in enumerate(lens):
            name = '%s%03d' % (split, v + 1)
            fd = os.path.join(root, 'raw_datasets', 'UCSDped2', split, name)
            od = os.path.join(root, 'optical_flow', 'UCSDped2', split, name)
            os.makedirs(fd, exist_ok=True), os.makedirs(od, exist_ok=True)
            if split == 'Test':
                gd = os.path.join(root, 'raw_datasets', 'UCSDped2', split, name + '_gt')
                os.makedirs(gd, exist_ok=True)
            for t in range(n):
                anomalous = split == 'Test' and t >= n // 2
                img, boxes, flow = _frame(t, 9.0 if anomalous else 2.0, seed=1000 * v + t)
                cv2.imwrite(os.path.join(fd, '%03d.tif' % (t + 1)), img)
                np.save(os.path.join(od, '%03d.npy' % (t + 1)), flow)
                bl.append(boxes)
                if split == 'Test':
                    cv2.imwrite(os.path.join(gd, '%03d.bmp' % (t + 1)), np.full((H, W), 255 if anomalous else 0, np.uint8))
    for split, bl in (('train', tr_boxes), ('test', te_boxes)):
        arr = np.empty(len(bl), dtype=object)
        for i, b in enumerate(bl):
            arr[i] = b
        np.save(os.path.join(root, 'raw_datasets', 'UCSDped2', 'bboxes_%s_obj_det_with_motion.npy' % split), arr)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = open(os.path.join(repo, 'config.cfg')).read()
    for k, v in (cfg_overrides or {}).items():
        lines = cfg.split('\n')
        hit = [i for i, l in enumerate(lines) if l.split('=')[0].strip() == k]
        assert hit, k
        for i in hit:
            lines[i] = '%s = %s' % (k, v)
        cfg = '\n'.join(lines)
    open(os.path.join(root, 'config.cfg'), 'w').write(cfg)
    return root
