// get_foreground on the device (reference: vad_datasets.py:70-93): crop every bounding box out of each frame of a frame stack and
// resize the crop to patch x patch with the arithmetic of cv2.resize(..., INTER_LINEAR) -- bit for bit, because the uint8 cubes
// this produces are an integer path of the pipeline.
//
// What cv2.resize does for these inputs (OpenCV 4.x imgproc/resize.cpp; pinned by tests/golden/crop_resize.npz, which
// tests/golden/make_crop_resize_golden.py wrote by calling cv2 itself):
//   * destination == source size: copy.
//   * source exactly 2x the destination in BOTH directions: INTER_LINEAR is replaced by the 2x2 box average
//     (uint8: (a + b + c + d + 2) >> 2; float: (((a + b) + c) + d) * 0.25f).
//   * otherwise, per destination column dx: fx = (float)((dx + 0.5) * scale_x - 0.5) (product and difference in double),
//     sx = floor(fx), fx -= sx; sx < 0 -> (sx, fx) = (0, 0); sx >= width - 1 -> (width - 1, 0).  Rows: same fy / sy, but the
//     weights are NOT reset at the border -- the two source rows are clipped to [0, height - 1] instead.
//     uint8 : coefficients as shorts cvRound(w * 2048); horizontal pass in int32 (sum of two products), vertical pass
//             ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
//     float : r = s0 * a0 + s1 * a1 per row, then r0 * b0 + r1 * b1, every product and sum rounded on its own (no FMA).
// One CTA per (box, frame), one thread per destination pixel, looping over the channels; frames are addressed through element
// strides so both cv2's HWC frames and the reference's CHW stacks can be passed as they are.
#include "common.h"

namespace {

struct CropParams {
    const void *frames;
    long long sT, sC, sH, sW;          // element strides of [T][C][H][W]
    int T, C, H, W;
    const int *boxes;                  // [N][4]: x_min, y_min, x_max, y_max (already ceil'ed on the host, as the reference does)
    int N, patch;
    void *out;                         // [N][T][C][patch][patch]
};

struct Tap {
    int i0, i1;
    float w0, w1;
    int q0, q1;                        // cvRound(w * 2048)
};

// coefficients of destination index d along one axis; clamp_weights: the horizontal rule (reset the weight at the border)
__device__ __forceinline__ Tap tap_of(int d, int ssize, int dsize, bool clamp_weights) {
    const double scale = (double)ssize / (double)dsize;
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);      // two roundings, as the host code (no DFMA contraction)
    int s = (int)floorf(f);
    f -= (float)s;
    Tap t;
    if (clamp_weights) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        t.i0 = s;
        t.i1 = min(s + 1, ssize - 1);
    } else {
        t.i0 = min(max(s, 0), ssize - 1);
        t.i1 = min(max(s + 1, 0), ssize - 1);
    }
    t.w0 = __fsub_rn(1.f, f);
    t.w1 = f;
    t.q0 = __float2int_rn(__fmul_rn(t.w0, 2048.f));
    t.q1 = __float2int_rn(__fmul_rn(t.w1, 2048.f));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(1024) k_crop_resize(const CropParams p) {
    vv_pdl_wait();
    const int n = blockIdx.x, t = blockIdx.y;
    const int ps = p.patch;
    const int dx = threadIdx.x, dy = threadIdx.y;
    const int x0 = p.boxes[4 * n + 0], y0 = p.boxes[4 * n + 1], x1 = p.boxes[4 * n + 2], y1 = p.boxes[4 * n + 3];
    const int sw = x1 - x0, sh = y1 - y0;             // validated on the host: 1 <= sw, sh and the box lies inside the frame
    const T *src = (const T *)p.frames + t * p.sT + y0 * p.sH + x0 * p.sW;
    T *dst = (T *)p.out + (((long long)n * p.T + t) * p.C) * ps * ps + dy * ps + dx;
    const bool copy = sw == ps && sh == ps, area = sw == 2 * ps && sh == 2 * ps;
    Tap tx, ty;
    if (!copy && !area) {
        tx = tap_of(dx, sw, ps, true);
        ty = tap_of(dy, sh, ps, false);
    }
    for (int c = 0; c < p.C; c++) {
        const T *s = src + c * p.sC;
        T v;
        if (copy) {
            v = s[dy * p.sH + dx * p.sW];
        } else if (area) {
            const T a = s[(2 * dy) * p.sH + (2 * dx) * p.sW], b = s[(2 * dy) * p.sH + (2 * dx + 1) * p.sW];
            const T cc = s[(2 * dy + 1) * p.sH + (2 * dx) * p.sW], d = s[(2 * dy + 1) * p.sH + (2 * dx + 1) * p.sW];
            if constexpr (sizeof(T) == 1) v = (T)(((int)a + (int)b + (int)cc + (int)d + 2) >> 2);
            else v = (T)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn((float)a, (float)b), (float)cc), (float)d), 0.25f);
        } else {
            const T a = s[ty.i0 * p.sH + tx.i0 * p.sW], b = s[ty.i0 * p.sH + tx.i1 * p.sW];
            const T cc = s[ty.i1 * p.sH + tx.i0 * p.sW], d = s[ty.i1 * p.sH + tx.i1 * p.sW];
            if constexpr (sizeof(T) == 1) {
                const int r0 = (int)a * tx.q0 + (int)b * tx.q1, r1 = (int)cc * tx.q0 + (int)d * tx.q1;
                const int o = (((ty.q0 * (r0 >> 4)) >> 16) + ((ty.q1 * (r1 >> 4)) >> 16) + 2) >> 2;
                v = (T)min(max(o, 0), 255);
            } else {
                const float r0 = __fadd_rn(__fmul_rn((float)a, tx.w0), __fmul_rn((float)b, tx.w1));
                const float r1 = __fadd_rn(__fmul_rn((float)cc, tx.w0), __fmul_rn((float)d, tx.w1));
                v = (T)__fadd_rn(__fmul_rn(r0, ty.w0), __fmul_rn(r1, ty.w1));
            }
        }
        dst[(long long)c * ps * ps] = v;
    }
}

}  // namespace

extern "C" int vecvad_crop_resize(const void *frames, int is_f32, int n_frames, int channels, int height, int width, int64_t stride_t,
                                  int64_t stride_c, int64_t stride_h, int64_t stride_w, const int32_t *boxes, int n_boxes, int patch,
                                  void *out, vecvad_stream stream) {
    VV_REQUIRE(frames && boxes && out, "crop_resize: null pointer");
    VV_REQUIRE(n_frames >= 1 && channels >= 1 && height >= 1 && width >= 1, "crop_resize: bad frame shape");
    VV_REQUIRE(patch >= 1 && patch <= 32, "crop_resize: patch %d not in [1, 32]", patch);
    if (n_boxes == 0) return 0;
    VV_REQUIRE(n_boxes > 0, "crop_resize: negative box count");
    CropParams p;
    p.frames = frames; p.sT = stride_t; p.sC = stride_c; p.sH = stride_h; p.sW = stride_w;
    p.T = n_frames; p.C = channels; p.H = height; p.W = width;
    p.boxes = boxes; p.N = n_boxes; p.patch = patch; p.out = out;
    const dim3 grid(n_boxes, n_frames), block(patch, patch);
    cudaError_t e = is_f32 ? vv_launch(k_crop_resize<float>, grid, block, 0, (cudaStream_t)stream, p)
                           : vv_launch(k_crop_resize<unsigned char>, grid, block, 0, (cudaStream_t)stream, p);
    VV_CK(e);
    VV_CKL();
    return 0;
}
