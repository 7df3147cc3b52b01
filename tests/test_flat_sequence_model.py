"""CPU model of the index arithmetic behind the flattened-sequence tiles (vec_vad_b200/csrc/igemm_flat.cu, wgrad_flat.cu).

The kernels never form an im2col matrix: a TMA box that is W + 1 pixels wide (the out-of-range column arrives as zeros) lands
in shared memory as a sequence in which image rows sit P = W + 1 positions apart, and every 3x3 tap is a constant offset
dy * P + dx into that sequence.  This file replays exactly that data movement in numpy -- same tile split (128 consecutive
positions of one image), same box (first row r_lo - 1, `rows` rows), same start rows (q0, g0), same guard row -- and checks it
against the op the reference calls (`nn.Conv2d(…, 3, padding=1)`, model/unet.py:10,13) and its autograd weight gradient, plus
the bounds the kernels rely on (no read before the one-row guard, none past the box for the rows that are read back).
No GPU needed: this is the host-side logic of those kernels.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

BM = 128  # positions per tile (TMEM lanes)


def box_rows(P):
    """image rows per activation box: rows holding 128 consecutive positions (<= 127 // P + 2) + one halo row either side
    (igemm_flat.cu: fp.rows; wgrad_flat.cu: a_rows = g_rows + 2)"""
    return (BM - 1) // P + 4


def load_box(img, y0, rows, P):
    """what the TMA box load leaves in shared memory: rows y0 .. y0+rows-1 of `img` [H, W, C], W + 1 wide, out-of-range = 0,
    flattened to [rows * P, C]"""
    H, W, C = img.shape
    box = np.zeros((rows, P, C), dtype=img.dtype)
    for r in range(rows):
        y = y0 + r
        if 0 <= y < H:
            box[r, :W] = img[y]
    return box.reshape(rows * P, C)


def read_rows(box, start, n):
    """n consecutive sequence rows from `start`; row -1 is the zeroed guard in front of the box (never further back)"""
    assert start >= -1, start
    out = np.zeros((n, box.shape[1]), dtype=box.dtype)
    lo = max(start, 0)
    hi = min(start + n, box.shape[0])
    out[lo - start:hi - start] = box[lo:hi]
    return out, start + n - 1


@pytest.mark.parametrize('H,W', [(32, 32), (16, 16), (64, 64), (8, 8), (20, 24), (4, 40)])
def test_flattened_conv_tiles_equal_conv2d(H, W):
    rng = np.random.default_rng(H * 100 + W)
    C, N = 4, 3
    x = rng.standard_normal((H, W, C))
    w = rng.standard_normal((N, C, 3, 3))
    want = F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w), None, padding=1)[0].permute(1, 2, 0).numpy()
    P = W + 1
    L = H * P
    tpi = (L + BM - 1) // BM
    rows = box_rows(P)
    got = np.full((H, W, N), np.nan)
    covered = np.zeros((H, W), dtype=int)
    for tt in range(tpi):
        r_lo = (tt * BM) // P                               # first image row with a position in this tile
        box = load_box(x, r_lo - 1, rows, P)                # box row 0 = image row r_lo - 1
        q0 = tt * BM - r_lo * P + P                         # box row of the tile's first position
        assert P <= q0 <= 2 * P - 1
        acc = np.zeros((BM, N))
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                a, last = read_rows(box, q0 + dy * P + dx, BM)      # the tap's A operand: same box, shifted start
                assert last < rows * P                               # never past the box
                acc += a @ w[:, :, dy + 1, dx + 1].T
        for lane in range(BM):                              # epilogue: position -> pixel; separator column / past the image discarded
            pos = tt * BM + lane
            y, xx = divmod(pos, P)
            if xx < W and y < H:
                got[y, xx] = acc[lane]
                covered[y, xx] += 1
    assert np.all(covered == 1)                             # every pixel produced by exactly one tile lane
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('H,W', [(32, 32), (16, 16), (64, 64), (8, 8), (20, 24)])
def test_flattened_wgrad_tiles_equal_autograd(H, W):
    rng = np.random.default_rng(H * 7 + W)
    C, N = 3, 5
    x = rng.standard_normal((H, W, C))
    go = rng.standard_normal((H, W, N))
    wt = torch.zeros(N, C, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], wt, None, padding=1).backward(torch.from_numpy(go).permute(2, 0, 1)[None])
    want = wt.grad.numpy()                                  # [N, C, 3, 3]
    P = W + 1
    L = H * P
    tpi = (L + BM - 1) // BM
    g_rows = (BM - 1) // P + 2
    a_rows = g_rows + 2
    assert a_rows == box_rows(P)
    dw = np.zeros((3, 3, N, C))                             # [dy+1][dx+1][n][k]
    for tt in range(tpi):
        r_lo = (tt * BM) // P
        abox = load_box(x, r_lo - 1, a_rows, P)             # activation box row 0 = image row r_lo - 1
        gbox = load_box(go, r_lo, g_rows, P)                # gradient box row 0 = image row r_lo
        g0 = tt * BM - r_lo * P
        assert 0 <= g0 <= P - 1
        for dyb in range(3):                                # M block dyb: activation rows from g0 + dyb * P  (dy = dyb - 1)
            a, last = read_rows(abox, g0 + dyb * P, BM)
            assert last < a_rows * P
            for j in range(3):                              # N block j: gradient rows from g0 - 1 + j  (dx = 1 - j)
                g, lastg = read_rows(gbox, g0 - 1 + j, BM)
                assert lastg < g_rows * P
                dw[dyb, 2 - j] += g.T @ a                   # contraction over the tile's 128 positions
        # the fourth M block (dy = +2) is computed and never read back; its reads stay inside the stage (activation + gradient box)
        assert g0 + 3 * P + BM - 1 < (a_rows + g_rows) * P + P
    got = dw.transpose(2, 3, 0, 1)                          # -> [n][k][dy][dx]
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


def test_tile_and_box_geometry_of_the_bench_shapes():
    """the numbers quoted in DESIGN.md section 3 for 32x32: 9 tiles per image, 7-row boxes, 89 % useful rows, 2.0x activation bytes
    (ncu: 209.7 MB through the SM<-L2 crossbar for 100.7 MB of input, profiles/r01_v5_step.txt)"""
    W = H = 32
    P = W + 1
    tpi = (H * P + BM - 1) // BM
    assert (P, tpi, box_rows(P)) == (33, 9, 7)
    assert abs(H * W / (tpi * BM) - 0.889) < 1e-3                       # useful MMA rows
    assert abs((box_rows(P) * P * tpi) / (H * W) - 2.03) < 0.01         # box bytes (separator column included) per input byte
