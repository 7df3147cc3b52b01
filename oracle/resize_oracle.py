"""CPU oracle for get_foreground's crop + resize (TEST INFRASTRUCTURE ONLY; never imported by vec_vad_b200).

Restates, in numpy integer / float32 arithmetic, what ``cv2.resize(crop, (patch, patch))`` -- the call inside the reference's
``get_foreground`` (vad_datasets.py:70-93) -- computes with its default INTER_LINEAR for the two element types the pipeline feeds
it (uint8 frames, float32 optical flow).  The algorithm lives in a third-party dependency that is not vendored under
/root/reference: OpenCV (README pins no version; this container has opencv-python 4.13.0).  Pinned by
tests/golden/crop_resize.npz, which tests/golden/make_crop_resize_golden.py wrote by calling cv2.resize itself on seeded crops
(tests/test_crop_resize.py holds this file to it bit for bit, on CPU).

OpenCV rules restated (modules/imgproc/src/resize.cpp, 4.x):
  * dsize == ssize: copy.
  * INTER_LINEAR with an exact 2x decimation in both directions is replaced by INTER_AREA's 2x2 box filter.
  * column dx: fx = (float)((dx + 0.5) * scale - 0.5) (double product / difference), sx = floor(fx), fx -= sx;
    sx < 0 -> sx = 0, fx = 0; sx >= w - 1 -> sx = w - 1, fx = 0.  Rows use the same fy / sy but keep the weights and clip the
    two row indices to [0, h - 1] instead.
  * uint8: weights as cvRound(w * 2048) shorts; rows r = s0 * a0 + s1 * a1 (int32); out = (((b0 * (r0 >> 4)) >> 16) +
    ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.   float32: r = s0 * a0 + s1 * a1, out = r0 * b0 + r1 * b1, every operation rounded.
"""
import numpy as np


def _coeffs(ssize, dsize, clamp_weights):
    scale = float(ssize) / dsize
    idx = np.empty(dsize, np.int64)
    w0, w1 = np.empty(dsize, np.float32), np.empty(dsize, np.float32)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_weights:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        idx[d], w0[d], w1[d] = s, np.float32(1.0) - f, f
    q0 = np.rint(w0 * np.float32(2048)).astype(np.int64)          # cvRound: half to even
    q1 = np.rint(w1 * np.float32(2048)).astype(np.int64)
    return idx, w0, w1, q0, q1


def resize_linear(src, ps):
    """src [h,w,C] uint8 or float32 -> [ps,ps,C], == cv2.resize(src, (ps, ps)) (INTER_LINEAR)."""
    h, w, _ = src.shape
    u8 = src.dtype == np.uint8
    if h == ps and w == ps:
        return src.copy()
    if h == 2 * ps and w == 2 * ps:
        if u8:
            s = src.astype(np.int64)
            return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
        return (((src[0::2, 0::2] + src[0::2, 1::2]) + src[1::2, 0::2] + src[1::2, 1::2]) * np.float32(0.25)).astype(np.float32)
    sx, ax0, ax1, qx0, qx1 = _coeffs(w, ps, True)
    sy, ay0, ay1, qy0, qy1 = _coeffs(h, ps, False)
    x1 = np.minimum(sx + 1, w - 1)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    out = np.empty((ps, ps, src.shape[2]), src.dtype)
    s = src.astype(np.int64) if u8 else src
    for dy in range(ps):
        if u8:
            r0 = s[y0[dy]][sx] * qx0[:, None] + s[y0[dy]][x1] * qx1[:, None]
            r1 = s[y1[dy]][sx] * qx0[:, None] + s[y1[dy]][x1] * qx1[:, None]
            out[dy] = np.clip((((qy0[dy] * (r0 >> 4)) >> 16) + ((qy1[dy] * (r1 >> 4)) >> 16) + 2) >> 2, 0, 255).astype(np.uint8)
        else:
            r0 = s[y0[dy]][sx] * ax0[:, None] + s[y0[dy]][x1] * ax1[:, None]
            r1 = s[y1[dy]][sx] * ax0[:, None] + s[y1[dy]][x1] * ax1[:, None]
            out[dy] = r0 * ay0[dy] + r1 * ay1[dy]
    return out


def get_foreground(img, bboxes, patch_size):
    """vad_datasets.py:70-93 with ``resize_linear`` in place of cv2.resize.  img [C,H,W] or [T,C,H,W]."""
    def one(frame, box):
        x_min, x_max = int(np.ceil(box[0])), int(np.ceil(box[2]))
        y_min, y_max = int(np.ceil(box[1])), int(np.ceil(box[3]))
        crop = np.ascontiguousarray(np.transpose(frame[:, y_min:y_max, x_min:x_max], [1, 2, 0]))
        return np.transpose(resize_linear(crop, patch_size), [2, 0, 1])
    if img.ndim == 3:
        return np.array([one(img, b) for b in bboxes])
    return np.array([np.array([one(img[j], b) for j in range(img.shape[0])]) for b in bboxes])
