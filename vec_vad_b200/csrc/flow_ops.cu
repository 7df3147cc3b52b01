// FlowNet2 native ops for sm_100a: correlation (cost volume), resample2d (bilinear warp), channelnorm,
// and the fused warp + difference + channel-norm used at every FlowNet2 stage boundary.
//
// Reference semantics restated from FlowNet2_src/models/components/ops:
//   correlation/src/correlation_cuda_kernel.cu:10-32 (pad + NHWC repack), :34-106 (forward), :108-290 (backward),
//   correlation/src/correlation_cuda.c:25-38 (output shape), resample2d/src/Resample2d_kernel.cu:20-66,69-186,
//   channelnorm/src/ChannelNorm_kernel.cu:19-51,54-81.
// Differences in HOW (not WHAT): inputs are read in place as NCHW with implicit zero padding (no padded NHWC
// scratch copies, no fill passes); the forward cost volume is register-tiled out of shared memory with the
// stride-2 displacement grid de-interleaved by column parity so every operand fetch is a 128-bit shared load.
#include "common.h"

long long vv_corr_tma_workspace(int batch, int channels, int h, int w);
int vv_launch_corr_tma(const float *in1, const float *in2, float *out, int batch, int channels, int h, int w, int pad, int md, int oc, int oh,
                       int ow, void *workspace, cudaStream_t st);

namespace {

__device__ __forceinline__ int cdiv_trunc(int a, int b) { return a / b; }   // C semantics (toward zero), as in the reference

struct CorrGeom {
    int B, C, H, W;          // input
    int pad, ksize, md, s1, s2;
    int kr, dr, D;           // kernel radius, displacement radius (in steps), grid size 2*dr+1
    int OC, OH, OW;
};

// ------------------------------------------------------------------------------------------------
// General forward: one thread per output pixel (x fastest -> coalesced), loops displacements, window, channels.
// ------------------------------------------------------------------------------------------------
__global__ void k_corr_fwd_general(const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out, CorrGeom g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int n = blockIdx.z / g.OC, tc = blockIdx.z % g.OC;
    if (x >= g.OW) return;
    const int tj = tc / g.D - g.dr, ti = tc % g.D - g.dr;
    // centre of the patch in UNPADDED coordinates (reference works in padded ones: y1 = y*s1 + md + kr)
    const int y1 = y * g.s1 + g.md + g.kr - g.pad, x1 = x * g.s1 + g.md + g.kr - g.pad;
    const int y2 = y1 + tj * g.s2, x2 = x1 + ti * g.s2;
    const long long HW = (long long)g.H * g.W;
    const float *a = in1 + (long long)n * g.C * HW, *b = in2 + (long long)n * g.C * HW;
    float acc = 0.f;
    for (int j = -g.kr; j <= g.kr; j++) {
        const int ya = y1 + j, yb = y2 + j;
        if (ya < 0 || ya >= g.H || yb < 0 || yb >= g.H) continue;
        for (int i = -g.kr; i <= g.kr; i++) {
            const int xa = x1 + i, xb = x2 + i;
            if (xa < 0 || xa >= g.W || xb < 0 || xb >= g.W) continue;
            const float *pa = a + (long long)ya * g.W + xa, *pb = b + (long long)yb * g.W + xb;
            for (int c = 0; c < g.C; c++) acc = fmaf(__ldg(pa + c * HW), __ldg(pb + c * HW), acc);
        }
    }
    out[(((long long)n * g.OC + tc) * g.OH + y) * g.OW + x] = acc / (float)(g.ksize * g.ksize * g.C);
}

// ------------------------------------------------------------------------------------------------
// FlowNetC forward (kernel 1, stride1 1, stride2 2, 21x21 displacements; FlowNetC.py:24-30).
// CTA = one output row y of one image, 64 consecutive x.  Thread = (tj, parity p, quad q): 4 outputs x = x0+p+2(4q+k)
// for all 21 ti -> 84 accumulators.  Channels are streamed through shared memory in chunks of CC with cp.async
// double buffering; each row of in2 is stored split by column parity so the 24 values a thread needs per channel
// (x + 2*ti, ti = -10..10, for its 4 x) are 6 consecutive float4.
// ------------------------------------------------------------------------------------------------
constexpr int FD = 21, FDR = 10, FTX = 64, FCC = 4;
constexpr int F2W = FTX + 4 * FDR;          // 104 columns of in2 per row
constexpr int F2H = F2W / 2;                // 52 per parity
constexpr int F_THREADS = 16 * FD;          // 336
constexpr int F_STAGE = FCC * (FTX + FD * F2W);   // floats per stage

__device__ __forceinline__ void cp_async4_zfill(float *dst, const float *src, bool ok) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    int sz = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}

__global__ void __launch_bounds__(F_THREADS) k_corr_fwd_flownetc(const float *__restrict__ in1, const float *__restrict__ in2,
                                                                 float *__restrict__ out, CorrGeom g) {
    extern __shared__ __align__(16) float sm[];   // [2][ f1: FCC x [2][32]  |  f2: FCC x FD x [2][52] ]
    const int x0 = blockIdx.x * FTX, y = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x;
    const int off = g.md - g.pad;                 // unpadded centre = output coordinate + off
    const int y1 = y + off, xb0 = x0 + off;       // in1 row / first column; in2 columns start at xb0 - 2*FDR
    const long long HW = (long long)g.H * g.W;
    const float *a = in1 + (long long)n * g.C * HW, *b = in2 + (long long)n * g.C * HW;

    auto stage = [&](int buf, int c0) {
        float *f1 = sm + buf * F_STAGE, *f2 = f1 + FCC * FTX;
        // in1: FCC x 64
        for (int i = t; i < FCC * FTX; i += F_THREADS) {
            int c = i / FTX, j = i - c * FTX;
            int xx = xb0 + j;
            bool ok = (c0 + c < g.C) && y1 >= 0 && y1 < g.H && xx >= 0 && xx < g.W;
            cp_async4_zfill(f1 + c * FTX + (j & 1) * (FTX / 2) + (j >> 1), ok ? a + (c0 + c) * HW + (long long)y1 * g.W + xx : a, ok);
        }
        // in2: FCC x 21 rows x 104
        for (int i = t; i < FCC * FD * F2W; i += F_THREADS) {
            int j = i % F2W, r = (i / F2W) % FD, c = i / (F2W * FD);
            int yy = y1 + (r - FDR) * 2, xx = xb0 - 2 * FDR + j;
            bool ok = (c0 + c < g.C) && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W;
            cp_async4_zfill(f2 + (c * FD + r) * F2W + (j & 1) * F2H + (j >> 1), ok ? b + (c0 + c) * HW + (long long)yy * g.W + xx : b, ok);
        }
        asm volatile("cp.async.commit_group;\n");
    };

    const int tj = t / 16, xg = t % 16, p = xg >> 3, q = xg & 7;
    float acc[4][FD];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < FD; i++) acc[k][i] = 0.f;

    const int nchunk = (g.C + FCC - 1) / FCC;
    stage(0, 0);
    for (int ch = 0; ch < nchunk; ch++) {
        if (ch + 1 < nchunk) {
            stage((ch + 1) & 1, (ch + 1) * FCC);
            asm volatile("cp.async.wait_group 1;\n");
        } else {
            asm volatile("cp.async.wait_group 0;\n");
        }
        __syncthreads();
        const float *f1 = sm + (ch & 1) * F_STAGE, *f2 = f1 + FCC * FTX;
#pragma unroll
        for (int c = 0; c < FCC; c++) {
            const float4 av = *reinterpret_cast<const float4 *>(f1 + c * FTX + p * (FTX / 2) + 4 * q);
            const float *row = f2 + (c * FD + tj) * F2W + p * F2H + 4 * q;
            float bv[24];
#pragma unroll
            for (int m = 0; m < 6; m++) {
                float4 v = *reinterpret_cast<const float4 *>(row + 4 * m);
                bv[4 * m] = v.x; bv[4 * m + 1] = v.y; bv[4 * m + 2] = v.z; bv[4 * m + 3] = v.w;
            }
            const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int i = 0; i < FD; i++) acc[k][i] = fmaf(aa[k], bv[k + i], acc[k][i]);
        }
        __syncthreads();
    }
    const float inv = 1.f / (float)g.C;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x0 + p + 2 * (4 * q + k);
        if (x < g.OW) {
#pragma unroll
            for (int i = 0; i < FD; i++) out[(((long long)n * g.OC + tj * FD + i) * g.OH + y) * g.OW + x] = acc[k][i] * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward (never executed by the VEC pipeline -- flow is computed without gradients, calc_optical_flow.py:56 --
// kept for API completeness).  One thread per input element, coalesced along x.
// ------------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void k_corr_bwd(const float *__restrict__ other, const float *__restrict__ gout, float *__restrict__ gin, CorrGeom g) {
    const int xu = blockIdx.x * blockDim.x + threadIdx.x;
    const int yu = blockIdx.y;
    const int n = blockIdx.z / g.C, c = blockIdx.z % g.C;
    if (xu >= g.W) return;
    const int y = yu + g.pad, x = xu + g.pad;                          // padded coordinates, as in the reference
    const long long HW = (long long)g.H * g.W, OHW = (long long)g.OH * g.OW;
    const float *oth = other + ((long long)n * g.C + c) * HW;
    const float *go = gout + (long long)n * g.OC * OHW;
    float acc = 0.f;
    for (int tc = 0; tc < g.OC; tc++) {
        const int i2 = (tc % g.D - g.dr) * g.s2, j2 = (tc / g.D - g.dr) * g.s2;
        const int sx = WHICH == 1 ? 0 : i2, sy = WHICH == 1 ? 0 : j2;
        int xmin = cdiv_trunc(x - g.kr - g.md - sx, g.s1), ymin = cdiv_trunc(y - g.kr - g.md - sy, g.s1);
        int xmax = cdiv_trunc(x + g.kr - g.md - sx, g.s1), ymax = cdiv_trunc(y + g.kr - g.md - sy, g.s1);
        if (xmax < 0 || ymax < 0 || xmin >= g.OW || ymin >= g.OH || xmin > xmax || ymin > ymax) continue;
        xmin = max(0, xmin); xmax = min(g.OW - 1, xmax);
        ymin = max(0, ymin); ymax = min(g.OH - 1, ymax);
        const int yo = WHICH == 1 ? yu + j2 : yu - j2, xo = WHICH == 1 ? xu + i2 : xu - i2;   // position in the other map (unpadded)
        if (yo < 0 || yo >= g.H || xo < 0 || xo >= g.W) continue;                             // zero padding
        const float v = __ldg(oth + (long long)yo * g.W + xo);
        float s = 0.f;
        for (int j = ymin; j <= ymax; j++)
            for (int i = xmin; i <= xmax; i++) s += __ldg(go + tc * OHW + (long long)j * g.OW + i);
        acc = fmaf(s, v, acc);
    }
    gin[((long long)n * g.C + c) * HW + (long long)yu * g.W + xu] = acc / (float)(g.ksize * g.ksize * g.C);
}

// ------------------------------------------------------------------------------------------------
// Resample2d: out[b,c,y,x] = bilinear(img[b,c], x + flow_x, y + flow_y); indices clamped to the border, weights from
// the unclamped fraction, products evaluated in double and accumulated into a float (Resample2d_kernel.cu:41-63).
// One thread per pixel: the flow and the four weights are computed once and reused for every channel.
// ------------------------------------------------------------------------------------------------
struct Bilin {
    int xL, xR, yT, yB;
    float alpha, beta;
};
__device__ __forceinline__ Bilin bilin(float dx, float dy, int x, int y, int cw, int chh) {
    Bilin r;
    float xf = (float)x + dx, yf = (float)y + dy;
    r.alpha = xf - floorf(xf);
    r.beta = yf - floorf(yf);
    r.xL = max(min((int)floorf(xf), cw - 1), 0);
    r.xR = max(min((int)(floorf(xf) + 1.f), cw - 1), 0);
    r.yT = max(min((int)floorf(yf), chh - 1), 0);
    r.yB = max(min((int)(floorf(yf) + 1.f), chh - 1), 0);
    return r;
}
__device__ __forceinline__ float bilin_sample(const float *im, int iw, const Bilin &w) {
    float val = 0.f;
    val += (1. - w.alpha) * (1. - w.beta) * __ldg(im + (long long)w.yT * iw + w.xL);
    val += (w.alpha) * (1. - w.beta) * __ldg(im + (long long)w.yT * iw + w.xR);
    val += (1. - w.alpha) * (w.beta) * __ldg(im + (long long)w.yB * iw + w.xL);
    val += (w.alpha) * (w.beta) * __ldg(im + (long long)w.yB * iw + w.xR);
    return val;
}

// img0 == nullptr: plain resample.  Otherwise also diff = img0 - warped (nullable) and norm = ||diff||_2 over channels (nullable).
__global__ void k_resample2d(const float *__restrict__ img, const float *__restrict__ flow, float *__restrict__ out,
                             const float *__restrict__ img0, float *__restrict__ diff, float *__restrict__ norm, int B, int C, int IH,
                             int IW, int OH, int OW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= OW) return;
    const long long OHW = (long long)OH * OW, IHW = (long long)IH * IW;
    const float dx = flow[((long long)b * 2 + 0) * OHW + (long long)y * OW + x];
    const float dy = flow[((long long)b * 2 + 1) * OHW + (long long)y * OW + x];
    const Bilin w = bilin(dx, dy, x, y, OW, OH);           // the reference clamps with the OUTPUT size (Resample2d_kernel.cu:48-51)
    float ss = 0.f;
    for (int c = 0; c < C; c++) {
        float v = bilin_sample(img + ((long long)b * C + c) * IHW, IW, w);
        const long long o = ((long long)b * C + c) * OHW + (long long)y * OW + x;
        if (out) out[o] = v;
        if (img0) {
            float d = img0[o] - v;
            if (diff) diff[o] = d;
            ss += d * d;
        }
    }
    if (norm) norm[(long long)b * OHW + (long long)y * OW + x] = sqrtf(ss);
}

// Fixed-channel-count variant (C = 1..3, what FlowNet2 uses: images and flows).  The four double weights of the reference's arithmetic --
// each tap's weight product in double, the running value rounded to float after every tap (Resample2d_kernel.cu:53-63) -- are formed
// ONCE per pixel instead of once per (channel, tap) and the channel loop is unrolled: same results bit for bit, 1.3x faster
// (profiles/r02_resample_variants.txt).  What bounds it on BASELINE.json configs[4]'s per-pixel N(0, 4 px) flow is the gather itself: a
// warp's 32 flow vectors point into ~20 different 128-byte lines per load instruction (L1 wavefronts), so even float arithmetic only
// reaches 0.30 of the HBM rate; staging the neighbourhood in shared memory was measured slower (same file) and dropped.
template <int C>
__global__ void __launch_bounds__(128) k_resample2d_c(const float *__restrict__ img, const float *__restrict__ flow, float *__restrict__ out,
                                                      const float *__restrict__ img0, float *__restrict__ diff, float *__restrict__ norm, int IH,
                                                      int IW, int OH, int OW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= OW) return;
    const long long OHW = (long long)OH * OW, IHW = (long long)IH * IW, pix = (long long)y * OW + x;
    const float dx = flow[((long long)b * 2 + 0) * OHW + pix], dy = flow[((long long)b * 2 + 1) * OHW + pix];
    const Bilin w = bilin(dx, dy, x, y, OW, OH);
    const long long oTL = (long long)w.yT * IW + w.xL, oTR = (long long)w.yT * IW + w.xR, oBL = (long long)w.yB * IW + w.xL,
                    oBR = (long long)w.yB * IW + w.xR;
    const double wTL = (1. - w.alpha) * (1. - w.beta), wTR = (w.alpha) * (1. - w.beta), wBL = (1. - w.alpha) * (w.beta), wBR = (w.alpha) * (w.beta);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const float *im = img + ((long long)b * C + c) * IHW;
        const float tl = __ldg(im + oTL), tr = __ldg(im + oTR), bl = __ldg(im + oBL), br = __ldg(im + oBR);
        float v = 0.f;
        v += wTL * tl; v += wTR * tr; v += wBL * bl; v += wBR * br;       // as bilin_sample: double product + sum, float after every tap
        const long long o = ((long long)b * C + c) * OHW + pix;
        if (out) out[o] = v;
        if (img0) {
            const float d = img0[o] - v;
            if (diff) diff[o] = d;
            ss += d * d;
        }
    }
    if (norm) norm[(long long)b * OHW + pix] = sqrtf(ss);
}

int launch_resample(const float *img, const float *flow, float *out, const float *img0, float *diff, float *norm, int B, int C, int IH, int IW,
                    int OH, int OW, cudaStream_t st) {
    static int generic = -1;                   // VECVAD_RESAMPLE_GENERIC=1: the any-channel-count kernel everywhere (bit-identical A/B baseline)
    if (generic < 0) { const char *g = getenv("VECVAD_RESAMPLE_GENERIC"); generic = (g && g[0] == '1') ? 1 : 0; }
    const dim3 grid(vv_cdiv(OW, 128), OH, B);
    if (!generic && C == 1) k_resample2d_c<1><<<grid, 128, 0, st>>>(img, flow, out, img0, diff, norm, IH, IW, OH, OW);
    else if (!generic && C == 2) k_resample2d_c<2><<<grid, 128, 0, st>>>(img, flow, out, img0, diff, norm, IH, IW, OH, OW);
    else if (!generic && C == 3) k_resample2d_c<3><<<grid, 128, 0, st>>>(img, flow, out, img0, diff, norm, IH, IW, OH, OW);
    else k_resample2d<<<grid, 128, 0, st>>>(img, flow, out, img0, diff, norm, B, C, IH, IW, OH, OW);
    VV_CKL();
    return 0;
}

// backward wrt the image: scatter-add (Resample2d_kernel.cu:69-116; weights use xf - int(xf), i.e. truncation, as there)
__global__ void k_resample2d_bwd_img(const float *__restrict__ flow, const float *__restrict__ gout, float *__restrict__ gimg, int B, int C,
                                     int IH, int IW, int OH, int OW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= OW) return;
    const long long OHW = (long long)OH * OW, IHW = (long long)IH * IW;
    const float dx = flow[((long long)b * 2 + 0) * OHW + (long long)y * OW + x];
    const float dy = flow[((long long)b * 2 + 1) * OHW + (long long)y * OW + x];
    const float xf = (float)x + dx, yf = (float)y + dy;
    const float alpha = xf - (float)(int)xf, beta = yf - (float)(int)yf;
    const int xL = max(min((int)floorf(xf), IW - 1), 0), xR = max(min((int)(floorf(xf) + 1.f), IW - 1), 0);
    const int yT = max(min((int)floorf(yf), IH - 1), 0), yB = max(min((int)(floorf(yf) + 1.f), IH - 1), 0);
    for (int c = 0; c < C; c++) {
        const float gv = gout[((long long)b * C + c) * OHW + (long long)y * OW + x];
        float *gi = gimg + ((long long)b * C + c) * IHW;
        atomicAdd(gi + (long long)yT * IW + xL, (1 - alpha) * (1 - beta) * gv);
        atomicAdd(gi + (long long)yT * IW + xR, (alpha) * (1 - beta) * gv);
        atomicAdd(gi + (long long)yB * IW + xL, (1 - alpha) * (beta) * gv);
        atomicAdd(gi + (long long)yB * IW + xR, (alpha) * (beta) * gv);
    }
}

// backward wrt the flow (Resample2d_kernel.cu:118-186, kernel_size 1)
__global__ void k_resample2d_bwd_flow(const float *__restrict__ img, const float *__restrict__ flow, const float *__restrict__ gout,
                                      float *__restrict__ gflow, int B, int C, int IH, int IW, int OH, int OW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= OW) return;
    const long long OHW = (long long)OH * OW, IHW = (long long)IH * IW;
    const float dx = flow[((long long)b * 2 + 0) * OHW + (long long)y * OW + x];
    const float dy = flow[((long long)b * 2 + 1) * OHW + (long long)y * OW + x];
    const float xf = (float)x + dx, yf = (float)y + dy;
    const int xL = max(min((int)floorf(xf), OW - 1), 0), xR = max(min((int)(floorf(xf) + 1.f), OW - 1), 0);
    const int yT = max(min((int)floorf(yf), OH - 1), 0), yB = max(min((int)(floorf(yf) + 1.f), OH - 1), 0);
    const float gx = 1.f - (yf - floorf(yf));     // weight used for d/d(flow_x)  (channel 0 branch)
    const float gy = 1.f - (xf - floorf(xf));     // weight used for d/d(flow_y)  (channel 1 branch)
    float ox = 0.f, oy = 0.f;
    for (int c = 0; c < C; c++) {
        const float gv = gout[((long long)b * C + c) * OHW + (long long)y * OW + x];
        const float *im = img + ((long long)b * C + c) * IHW;
        const float tl = __ldg(im + (long long)yT * IW + xL), tr = __ldg(im + (long long)yT * IW + xR);
        const float bl = __ldg(im + (long long)yB * IW + xL), br = __ldg(im + (long long)yB * IW + xR);
        ox += gx * gv * tr; ox -= gx * gv * tl; ox += (1 - gx) * gv * br; ox -= (1 - gx) * gv * bl;
        oy += gy * gv * bl; oy -= gy * gv * tl; oy += (1 - gy) * gv * br; oy -= (1 - gy) * gv * tr;
    }
    gflow[((long long)b * 2 + 0) * OHW + (long long)y * OW + x] = ox;
    gflow[((long long)b * 2 + 1) * OHW + (long long)y * OW + x] = oy;
}

// ---- ChannelNorm (ChannelNorm_kernel.cu:19-51, 54-81); norm_deg is ignored by the reference kernels too (always L2)
__global__ void k_channelnorm(const float *__restrict__ in, float *__restrict__ out, int C, long long HW) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= HW) return;
    float s = 0.f;
    for (int c = 0; c < C; c++) {
        float v = in[((long long)b * C + c) * HW + i];
        s += v * v;
    }
    out[(long long)b * HW + i] = sqrtf(s);
}
__global__ void k_channelnorm_bwd(const float *__restrict__ in, const float *__restrict__ out, const float *__restrict__ gout,
                                  float *__restrict__ gin, int C, long long HW) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= HW) return;
    const float g = gout[(long long)b * HW + i];
    const double den = (double)out[(long long)b * HW + i] + 1e-9;      // float + double literal, as in the reference
    for (int c = 0; c < C; c++) {
        const long long o = ((long long)b * C + c) * HW + i;
        gin[o] = (float)((double)(g * in[o]) / den);
    }
}

int corr_geom(CorrGeom &g, int batch, int channels, int in_h, int in_w, int pad, int ksize, int md, int s1, int s2) {
    VV_REQUIRE(batch >= 1 && channels >= 1 && in_h >= 1 && in_w >= 1, "correlation: bad tensor shape");
    VV_REQUIRE(pad >= 0 && ksize >= 1 && (ksize & 1) && md >= 0 && s1 >= 1 && s2 >= 1, "correlation: bad parameters");
    g.B = batch; g.C = channels; g.H = in_h; g.W = in_w;
    g.pad = pad; g.ksize = ksize; g.md = md; g.s1 = s1; g.s2 = s2;
    g.kr = (ksize - 1) / 2; g.dr = md / s2; g.D = 2 * g.dr + 1;
    const int border = g.kr + md;
    const int ph = in_h + 2 * pad, pw = in_w + 2 * pad;
    g.OC = g.D * g.D;
    g.OH = (ph - 2 * border + s1 - 1) / s1;      // ceil((float)(ph - 2*border) / s1), correlation_cuda.c:33-34
    g.OW = (pw - 2 * border + s1 - 1) / s1;
    VV_REQUIRE(ph - 2 * border > 0 && pw - 2 * border > 0, "correlation: empty output (input %dx%d, pad %d, border %d)", in_h, in_w, pad, border);
    return 0;
}

}  // namespace

extern "C" int vecvad_correlation_out_shape(int in_h, int in_w, int pad_size, int kernel_size, int max_displacement, int stride1,
                                            int stride2, int *out_c, int *out_h, int *out_w) {
    CorrGeom g;
    int r = corr_geom(g, 1, 1, in_h, in_w, pad_size, kernel_size, max_displacement, stride1, stride2);
    if (r) return r;
    if (out_c) *out_c = g.OC;
    if (out_h) *out_h = g.OH;
    if (out_w) *out_w = g.OW;
    return 0;
}

static bool corr_fast_params(const CorrGeom &g, int batch, int kernel_size, int stride1, int stride2) {
    return kernel_size == 1 && stride1 == 1 && stride2 == 2 && g.D == FD && g.OH <= 65535 && batch <= 32767;
}

extern "C" int vecvad_correlation_workspace_bytes(int batch, int channels, int in_h, int in_w, int pad_size, int kernel_size,
                                                  int max_displacement, int stride1, int stride2, int64_t *bytes) {
    VV_REQUIRE(bytes, "correlation_workspace_bytes: null pointer");
    CorrGeom g;
    int r = corr_geom(g, batch, channels, in_h, in_w, pad_size, kernel_size, max_displacement, stride1, stride2);
    if (r) return r;
    *bytes = (corr_fast_params(g, batch, kernel_size, stride1, stride2) && pad_size == max_displacement)
                 ? vv_corr_tma_workspace(batch, channels, in_h, in_w) : 0;
    return 0;
}

extern "C" int vecvad_correlation_forward(const float *in1, const float *in2, float *out, int batch, int channels, int in_h, int in_w,
                                          int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                          int corr_type_multiply, void *workspace, int64_t workspace_bytes, vecvad_stream stream) {
    (void)corr_type_multiply;   // ignored by the reference kernels as well (always a product)
    VV_REQUIRE(in1 && in2 && out, "correlation_forward: null pointer");
    CorrGeom g;
    int r = corr_geom(g, batch, channels, in_h, in_w, pad_size, kernel_size, max_displacement, stride1, stride2);
    if (r) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const bool fast = corr_fast_params(g, batch, kernel_size, stride1, stride2);
    const long long need = fast ? vv_corr_tma_workspace(batch, channels, in_h, in_w) : 0;
    if (fast && pad_size == max_displacement && workspace && need > 0 && workspace_bytes >= need)
        return vv_launch_corr_tma(in1, in2, out, batch, channels, in_h, in_w, pad_size, max_displacement, g.OC, g.OH, g.OW, workspace, st);
    if (fast) {                  // FlowNetC parameters without scratch: cp.async-staged kernel reading the inputs in place
        const size_t smem = 2 * F_STAGE * sizeof(float);
        static bool attr = false;
        if (!attr) {
            VV_CK(cudaFuncSetAttribute(k_corr_fwd_flownetc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        k_corr_fwd_flownetc<<<dim3(vv_cdiv(g.OW, FTX), g.OH, batch), F_THREADS, smem, st>>>(in1, in2, out, g);
    } else {
        VV_REQUIRE((long long)batch * g.OC <= 65535 && g.OH <= 65535, "correlation_forward: batch*displacements too large for the general kernel");
        k_corr_fwd_general<<<dim3(vv_cdiv(g.OW, 128), g.OH, batch * g.OC), 128, 0, st>>>(in1, in2, out, g);
    }
    VV_CKL();
    return 0;
}

extern "C" int vecvad_correlation_backward(const float *in1, const float *in2, const float *grad_out, float *grad_in1, float *grad_in2,
                                           int batch, int channels, int in_h, int in_w, int pad_size, int kernel_size,
                                           int max_displacement, int stride1, int stride2, int corr_type_multiply, vecvad_stream stream) {
    (void)corr_type_multiply;
    VV_REQUIRE(in1 && in2 && grad_out && grad_in1 && grad_in2, "correlation_backward: null pointer");
    VV_REQUIRE(stride1 == 1, "correlation_backward: stride1 != 1 writes out of bounds in the reference (correlation_cuda_kernel.cu:121-122,195); unsupported");
    CorrGeom g;
    int r = corr_geom(g, batch, channels, in_h, in_w, pad_size, kernel_size, max_displacement, stride1, stride2);
    if (r) return r;
    VV_REQUIRE((long long)batch * channels <= 65535 && in_h <= 65535, "correlation_backward: batch*channels too large");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(vv_cdiv(in_w, 128), in_h, batch * channels);
    k_corr_bwd<1><<<grid, 128, 0, st>>>(in2, grad_out, grad_in1, g);
    VV_CKL();
    k_corr_bwd<2><<<grid, 128, 0, st>>>(in1, grad_out, grad_in2, g);
    VV_CKL();
    return 0;
}

extern "C" int vecvad_resample2d_forward(const float *img, const float *flow, float *out, int batch, int channels, int img_h, int img_w,
                                         int out_h, int out_w, int kernel_size, vecvad_stream stream) {
    VV_REQUIRE(img && flow && out, "resample2d_forward: null pointer");
    VV_REQUIRE(kernel_size == 1, "resample2d: only kernel_size 1 is supported (the only value any caller uses: modules/resample2d.py:8)");
    VV_REQUIRE(out_h <= img_h && out_w <= img_w, "resample2d: flow larger than the image reads out of bounds in the reference; unsupported");
    VV_REQUIRE(out_h <= 65535 && batch <= 65535, "resample2d: tensor too large");
    return launch_resample(img, flow, out, nullptr, nullptr, nullptr, batch, channels, img_h, img_w, out_h, out_w, (cudaStream_t)stream);
}

extern "C" int vecvad_resample2d_backward(const float *img, const float *flow, const float *grad_out, float *grad_img, float *grad_flow,
                                          int batch, int channels, int img_h, int img_w, int out_h, int out_w, int kernel_size,
                                          vecvad_stream stream) {
    VV_REQUIRE(img && flow && grad_out && grad_img && grad_flow, "resample2d_backward: null pointer");
    VV_REQUIRE(kernel_size == 1, "resample2d: only kernel_size 1 is supported");
    VV_REQUIRE(out_h <= img_h && out_w <= img_w && out_h <= 65535 && batch <= 65535, "resample2d_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    VV_CK(cudaMemsetAsync(grad_img, 0, (size_t)batch * channels * img_h * img_w * sizeof(float), st));
    dim3 grid(vv_cdiv(out_w, 128), out_h, batch);
    k_resample2d_bwd_img<<<grid, 128, 0, st>>>(flow, grad_out, grad_img, batch, channels, img_h, img_w, out_h, out_w);
    VV_CKL();
    k_resample2d_bwd_flow<<<grid, 128, 0, st>>>(img, flow, grad_out, grad_flow, batch, channels, img_h, img_w, out_h, out_w);
    VV_CKL();
    return 0;
}

extern "C" int vecvad_channelnorm_forward(const float *in, float *out, int batch, int channels, int h, int w, int norm_deg,
                                          vecvad_stream stream) {
    (void)norm_deg;
    VV_REQUIRE(in && out && batch >= 1 && batch <= 65535, "channelnorm_forward: bad arguments");
    const long long HW = (long long)h * w;
    k_channelnorm<<<dim3(vv_cdiv(HW, 256), batch), 256, 0, (cudaStream_t)stream>>>(in, out, channels, HW);
    VV_CKL();
    return 0;
}

extern "C" int vecvad_channelnorm_backward(const float *in, const float *out, const float *grad_out, float *grad_in, int batch,
                                           int channels, int h, int w, int norm_deg, vecvad_stream stream) {
    (void)norm_deg;
    VV_REQUIRE(in && out && grad_out && grad_in && batch >= 1 && batch <= 65535, "channelnorm_backward: bad arguments");
    const long long HW = (long long)h * w;
    k_channelnorm_bwd<<<dim3(vv_cdiv(HW, 256), batch), 256, 0, (cudaStream_t)stream>>>(in, out, grad_out, grad_in, channels, HW);
    VV_CKL();
    return 0;
}

extern "C" int vecvad_warp_diff_norm(const float *img0, const float *img1, const float *flow, float *warped, float *diff, float *norm,
                                     int batch, int channels, int h, int w, vecvad_stream stream) {
    VV_REQUIRE(img0 && img1 && flow, "warp_diff_norm: null pointer");
    VV_REQUIRE(h <= 65535 && batch <= 65535, "warp_diff_norm: tensor too large");
    return launch_resample(img1, flow, warped, img0, diff, norm, batch, channels, h, w, h, w, (cudaStream_t)stream);
}
