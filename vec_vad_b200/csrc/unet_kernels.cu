// Bandwidth-bound kernels of the completion-UNet set: input staging, weight re-layout, BatchNorm
// (train/eval) + ReLU (+2x2 max-pool) apply, BatchNorm backward, max-pool backward, 1x1 output conv
// with fused squared-error / MSE gradient, Adam.  All operate on grouped NHWC fp32 tensors
// [G][B*H*W][ld] (blockIdx.z / .y = UNet index g).
//
// Reference semantics restated (reference file:line in each kernel's comment); PyTorch layer
// definitions at model/unet.py:9-16 (conv+BN+ReLU), :38 (MaxPool2d(2)), :54 (ConvTranspose2d), :66 (1x1 conv).
#include <cuda_fp16.h>

#include "unet_kernels.h"

namespace {

// element-typed 4-wide access: H = fp16 storage (8 bytes), else fp32 (16 bytes); idx in elements
__device__ __forceinline__ uint2 pack_h4(const float4 &v) {
    const float m = 65504.f;                     // saturate instead of producing infinities
    __half2 lo = __floats2half2_rn(fminf(fmaxf(v.x, -m), m), fminf(fmaxf(v.y, -m), m));
    __half2 hi = __floats2half2_rn(fminf(fmaxf(v.z, -m), m), fminf(fmaxf(v.w, -m), m));
    uint2 o;
    o.x = *reinterpret_cast<unsigned *>(&lo);
    o.y = *reinterpret_cast<unsigned *>(&hi);
    return o;
}
template <bool H>
__device__ __forceinline__ float4 ld4(const void *base, long long idx) {
    if (H) {
        const uint2 r = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(base) + idx);
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(base) + idx);
}
template <bool H>
__device__ __forceinline__ void st4(void *base, long long idx, const float4 &v) {
    if (H) *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(base) + idx) = pack_h4(v);
    else *reinterpret_cast<float4 *>(reinterpret_cast<float *>(base) + idx) = v;
}

// eight consecutive elements: one 16-byte access for fp16, two for fp32
template <bool H>
__device__ __forceinline__ void ld8(const void *base, long long idx, float *v) {
    if (H) {
        const uint4 r = *reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(base) + idx);
        const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    } else {
        const float4 a = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(base) + idx);
        const float4 b = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(base) + idx + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
template <bool H>
__device__ __forceinline__ void st8(void *base, long long idx, const float *v) {
    if (H) {
        const uint2 lo = pack_h4(make_float4(v[0], v[1], v[2], v[3])), hi = pack_h4(make_float4(v[4], v[5], v[6], v[7]));
        *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(base) + idx) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    } else {
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(base) + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__device__ __forceinline__ void st1(void *base, long long idx, float v, int h) {
    if (h) reinterpret_cast<__half *>(base)[idx] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else reinterpret_cast<float *>(base)[idx] = v;
}

__device__ __forceinline__ void pix3(int m, int H, int W, int &b, int &y, int &x) {
    x = m % W;
    int t = m / W;
    y = t % H;
    b = t / H;
}

// ---- input staging: x NCHW [B, 3*T, S, S] -> X0 [G][B*S*S][cinp] with frame erase[g] dropped or zeroed (model/unet.py:179-183)
template <bool H>
__global__ void k_prep_input(const float *__restrict__ x, void *__restrict__ X0, int B, int T, int S, int cinp, int padding,
                             VvIntG erase) {
    vv_pdl_wait();
    const int g = blockIdx.y;
    const int M = B * S * S;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int pix = m % (S * S), b = m / (S * S);
    const int e = erase.v[g];
    const float *xb = x + (long long)b * 3 * T * S * S + pix;
    const long long dst = ((long long)g * M + m) * cinp;
    const int creal = padding ? 3 * T : 3 * (T - 1);
    for (int c4 = 0; c4 < cinp; c4 += 8) {              // cinp is a multiple of 32: 16-byte stores of eight channels
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int c = c4 + j;
            float val = 0.f;
            if (c < creal) {
                int src = padding ? c : (c < 3 * e ? c : c + 3);
                bool erased = padding && (c >= 3 * e) && (c < 3 * e + 3);
                if (!erased) val = __ldg(xb + (long long)src * S * S);
            }
            v[j] = val;
        }
        st8<H>(X0, dst + c4, v);
    }
}

// ---- weight re-layout, once per step.  Conv2d weight [N][C][3][3] -> Wf[t][N][Cp] (forward B operand, K contiguous)
//      and Wd[t][C][N] (dgrad B operand); bias / gamma / beta gathered into a uniform-stride vector block [3][N].
__global__ void k_prep_conv_w(const float *__restrict__ params, VvIntG slot, long long slot_stride, long long w_off, long long b_off,
                              long long g_off, long long beta_off, int N, int C, int Cp, void *__restrict__ Wf, long long wf_gs,
                              void *__restrict__ Wd, long long wd_gs, int w_f16, float *__restrict__ vec, long long vec_gs) {
    vv_pdl_wait();
    const int g = blockIdx.y;
    const float *P = params + slot.v[g] * slot_stride;
    const int total = 9 * N * Cp;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        int c = i % Cp;
        int n = (i / Cp) % N;
        int t = i / (Cp * N);
        float v = (c < C) ? P[w_off + ((long long)n * C + c) * 9 + t] : 0.f;
        st1(Wf, g * wf_gs + i, v, w_f16);
        if (Wd && c < C) st1(Wd, g * wd_gs + ((long long)t * C + c) * N + n, v, w_f16);
    }
    if (blockIdx.x == 0 && vec) {
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            vec[g * vec_gs + n] = P[b_off + n];
            vec[g * vec_gs + N + n] = P[g_off + n];
            vec[g * vec_gs + 2 * N + n] = P[beta_off + n];
        }
    }
}

// Same re-layout, tiled through shared memory: block = 32 output channels x 32 input channels x 9 taps, so the reads
// (288 contiguous floats per output channel) and both writes (128-byte runs along c for Wf, along n for Wd) are coalesced.
__global__ void __launch_bounds__(256) k_prep_conv_w_tiled(const float *__restrict__ params, VvIntG slot, long long slot_stride,
                                                           long long w_off, long long b_off, long long g_off, long long beta_off, int N,
                                                           int C, int Cp, void *__restrict__ Wf, long long wf_gs, void *__restrict__ Wd,
                                                           long long wd_gs, int w_f16, float *__restrict__ vec, long long vec_gs) {
    vv_pdl_wait();
    __shared__ float tile[32][32 * 9 + 1];
    const int g = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *P = params + slot.v[g] * slot_stride;
    const int cw = min(32, C - c0);                       // real input channels in this tile (<= 0: all padding)
    for (int i = threadIdx.x; i < 32 * 288; i += 256) {
        const int n = i / 288, r = i - n * 288;           // r = c_local * 9 + t
        tile[n][r] = (r < cw * 9) ? P[w_off + ((long long)(n0 + n) * C + c0) * 9 + r] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
        const int c = i & 31, n = (i >> 5) & 31, t = i >> 10;
        st1(Wf, g * wf_gs + ((long long)t * N + n0 + n) * Cp + c0 + c, tile[n][c * 9 + t], w_f16);
    }
    if (Wd) {
        for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
            const int n = i & 31, c = (i >> 5) & 31, t = i >> 10;
            if (c < cw) st1(Wd, g * wd_gs + ((long long)t * C + c0 + c) * N + n0 + n, tile[n][c * 9 + t], w_f16);
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && vec) {
        for (int n = threadIdx.x; n < N; n += 256) {
            vec[g * vec_gs + n] = P[b_off + n];
            vec[g * vec_gs + N + n] = P[g_off + n];
            vec[g * vec_gs + 2 * N + n] = P[beta_off + n];
        }
    }
}

// one launch for every conv unit of the net: block -> (unit, 32 x 32 channel tile)
__device__ __forceinline__ int find_unit(const VvPrepAll &all, int blk) {
    int u = 0;
    while (u + 1 < all.n && blk >= all.u[u + 1].blk0) u++;
    return u;
}
__global__ void __launch_bounds__(256) k_prep_conv_w_all(const float *__restrict__ params, VvIntG slot, long long slot_stride,
                                                         const VvPrepAll all) {
    vv_pdl_wait();
    __shared__ float tile[32][32 * 9 + 1];
    const int ui = find_unit(all, blockIdx.x);
    const VvPrepUnit &U = all.u[ui];
    const int lb = blockIdx.x - U.blk0, nbx = U.N / 32;
    const int g = blockIdx.z, n0 = (lb % nbx) * 32, c0 = (lb / nbx) * 32;
    const float *P = params + slot.v[g] * slot_stride;
    const int N = U.N, C = U.C, Cp = U.Cp;
    const int cw = min(32, C - c0);
    for (int i = threadIdx.x; i < 32 * 288; i += 256) {
        const int n = i / 288, r = i - n * 288;
        tile[n][r] = (r < cw * 9) ? P[U.w_off + ((long long)(n0 + n) * C + c0) * 9 + r] : 0.f;
    }
    __syncthreads();
    // four consecutive elements per store (8 bytes fp16 / 16 bytes fp32): [tap][n][c0 + 4q ..] and, transposed, [tap][c][n0 + 4q ..]
    for (int i = threadIdx.x; i < 9 * 32 * 8; i += 256) {
        const int q = i & 7, n = (i >> 3) & 31, t = i >> 8;
        const float4 v = make_float4(tile[n][(4 * q) * 9 + t], tile[n][(4 * q + 1) * 9 + t], tile[n][(4 * q + 2) * 9 + t], tile[n][(4 * q + 3) * 9 + t]);
        const long long o = g * U.wf_gs + ((long long)t * N + n0 + n) * Cp + c0 + 4 * q;
        if (all.w_f16) st4<true>(U.Wf, o, v); else st4<false>(U.Wf, o, v);
    }
    if (U.Wd) {
        for (int i = threadIdx.x; i < 9 * 32 * 8; i += 256) {
            const int q = i & 7, c = (i >> 3) & 31, t = i >> 8;
            if (c < cw) {
                const float4 v = make_float4(tile[4 * q][c * 9 + t], tile[4 * q + 1][c * 9 + t], tile[4 * q + 2][c * 9 + t], tile[4 * q + 3][c * 9 + t]);
                const long long o = g * U.wd_gs + ((long long)t * C + c0 + c) * N + n0 + 4 * q;
                if (all.w_f16) st4<true>(U.Wd, o, v); else st4<false>(U.Wd, o, v);
            }
        }
    }
    if (lb == 0) {
        for (int n = threadIdx.x; n < N; n += 256) {
            U.vec[g * U.vec_gs + n] = P[U.b_off + n];
            U.vec[g * U.vec_gs + N + n] = P[U.g_off + n];
            U.vec[g * U.vec_gs + 2 * N + n] = P[U.beta_off + n];
        }
    }
}
__global__ void __launch_bounds__(256) k_scatter_conv_wgrad_all(float *__restrict__ grads, VvIntG slot, long long slot_stride,
                                                                const VvPrepAll all) {
    vv_pdl_wait();
    __shared__ float tile[32][32 * 9 + 1];
    const int ui = find_unit(all, blockIdx.x);
    const VvPrepUnit &U = all.u[ui];
    const int lb = blockIdx.x - U.blk0, nbx = U.N / 32;
    const int g = blockIdx.z, n0 = (lb % nbx) * 32, c0 = (lb / nbx) * 32;
    const int N = U.N, C = U.C, Cp = U.Cp;
    const int cw = min(32, C - c0);
    if (cw <= 0) return;
    for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
        const int c = i & 31, n = (i >> 5) & 31, t = i >> 10;
        tile[n][c * 9 + t] = U.dWf[g * U.wf_gs + ((long long)t * N + n0 + n) * Cp + c0 + c];
    }
    __syncthreads();
    float *Gp = grads + slot.v[g] * slot_stride + U.w_off;
    for (int i = threadIdx.x; i < 32 * 288; i += 256) {
        const int n = i / 288, r = i - n * 288;
        if (r < cw * 9) Gp[((long long)(n0 + n) * C + c0) * 9 + r] = tile[n][r] * all.scale;
    }
}

// ConvTranspose2d(k3,s2,p1,op1) weight [Ci][Co][3][3] -> 2x2-tap "big" matrices over the 4 output phases:
//   out[2y+py, 2x+px, co] = sum_{sy,sx in {0,1}} in[y+sy, x+sx, :] . Wt[:, co, ky, kx],  ky = py+1-2sy, kx = px+1-2sx (if in 0..2)
//   Wbf[s][(p,co)][ci]  (forward, N = 4Co, Kt = Ci)      Wbd[s][ci][(p,co)]  (input gradient, N = Ci, Kt = 4Co)
__global__ void k_prep_ct_w(const float *__restrict__ params, VvIntG slot, long long slot_stride, long long w_off, long long b_off,
                            int Ci, int Co, void *__restrict__ Wbf, long long wf_gs, void *__restrict__ Wbd, long long wd_gs, int w_f16,
                            float *__restrict__ vec, long long vec_gs) {
    vv_pdl_wait();
    const int g = blockIdx.y;
    const float *P = params + slot.v[g] * slot_stride;
    const int total = 4 * 4 * Co * Ci;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        int ci = i % Ci;
        int co = (i / Ci) % Co;
        int ph = (i / (Ci * Co)) % 4;
        int s = i / (Ci * Co * 4);
        int ky = (ph >> 1) + 1 - 2 * (s >> 1), kx = (ph & 1) + 1 - 2 * (s & 1);
        float v = 0.f;
        if (ky >= 0 && kx >= 0) v = P[w_off + (((long long)ci * Co + co) * 3 + ky) * 3 + kx];
        st1(Wbf, g * wf_gs + i, v, w_f16);                                                     // [s][ph*Co+co][ci]
        st1(Wbd, g * wd_gs + ((long long)s * Ci + ci) * (4 * Co) + ph * Co + co, v, w_f16);    // [s][ci][ph*Co+co]
    }
    if (blockIdx.x == 0)
        for (int n = threadIdx.x; n < Co; n += blockDim.x) vec[g * vec_gs + n] = P[b_off + n];
}

// ---- BatchNorm2d (+ReLU, + optional 2x2 max-pool).  nn.BatchNorm2d defaults: eps 1e-5, momentum 0.1, biased variance for
//      normalisation, unbiased for running_var.  Z [G][M][C] raw conv output; stats [G][2][C] double sums from the conv epilogue.
__device__ __forceinline__ void bn_scale_shift(const VvBnApply &p, int g, int c, float &scale, float &shift, bool writer) {
    const float *vec = p.vec + g * p.vec_gs;
    float gamma = vec[p.C + c], beta = vec[2 * p.C + c];
    float mean, invstd;
    float *run = p.running + p.slot.v[g] * p.slot_stat_stride;
    if (p.training) {
        const double *st = p.stats + g * p.stats_gs;
        double inv_m = 1.0 / (double)p.M;
        double mu = st[c] * inv_m;
        double var = st[p.C + c] * inv_m - mu * mu;
        if (var < 0.0) var = 0.0;
        mean = (float)mu;
        invstd = (float)(1.0 / sqrt(var + 1e-5));
        if (writer) {
            float unb = (p.M > 1) ? (float)(var * ((double)p.M / (double)(p.M - 1))) : (float)var;
            run[p.rm_off + c] = 0.9f * run[p.rm_off + c] + 0.1f * mean;
            run[p.rv_off + c] = 0.9f * run[p.rv_off + c] + 0.1f * unb;
        }
    } else {
        mean = run[p.rm_off + c];
        invstd = 1.0f / sqrtf(run[p.rv_off + c] + 1e-5f);
    }
    scale = gamma * invstd;
    shift = beta - mean * scale;
    if (writer && p.save) {
        float *sv = p.save + g * p.save_gs;
        sv[c] = scale; sv[p.C + c] = shift; sv[2 * p.C + c] = mean; sv[3 * p.C + c] = invstd;
    }
}

// ---- element-typed row access for the bandwidth-bound passes: a thread owns CH = 4 consecutive channels of a pixel row (16 bytes of
//      fp32, 8 bytes of fp16) and keeps UNR rows in flight.  Measured (profiles/r02_bn_ncu.txt): 8 channels per thread (16-byte fp16
//      accesses) costs registers (116 / thread, 25 % occupancy) and ran slower; these passes need threads in flight, not wider loads.
template <bool H> struct RowVec { static constexpr int CH = 4; float v[CH]; };
template <bool H>
__device__ __forceinline__ RowVec<H> ldv(const void *base, long long idx) {
    RowVec<H> r;
    const float4 q = ld4<H>(base, idx);
    r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    return r;
}
template <bool H>
__device__ __forceinline__ void stv(void *base, long long idx, const RowVec<H> &r) {
    st4<H>(base, idx, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
}

// thread = (output pixel or pooled pixel, CH channels).  blockDim = (C/CH <= 64, rows).  H: Z, Y and P are fp16 (Y / P are the
// operands of the next contraction; rounded once, here, exactly as the tf32 path rounds them on their way into shared memory)
template <bool H>
__global__ void __launch_bounds__(256, 4) k_bn_apply(const VvBnApply p) {
    vv_pdl_wait();
    constexpr int CH = RowVec<H>::CH;
    constexpr int UNR = 4;                 // rows in flight per thread: these passes are bound by bytes in flight
    extern __shared__ float sm[];          // scale[C], shift[C]
    const int g = blockIdx.y;
    float *s_scale = sm, *s_shift = sm + p.C;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    for (int c = tid; c < p.C; c += nthr) {
        float sc, sh;
        bn_scale_shift(p, g, c, sc, sh, blockIdx.x == 0);
        s_scale[c] = sc; s_shift[c] = sh;
    }
    __syncthreads();
    const long long z0 = g * p.z_gs;
    const long long y0 = g * p.y_gs + p.y_coff;
    const int cg = p.C / CH;
    // rev: rows are walked from the END -- the conv tile that produced Z wrote its last rows most recently (still in L2), and the
    // rows written last here are the ones the next conv tile reads first
    auto RV = [&](long long r, long long n) -> long long { return p.rev ? n - 1 - r : r; };
    for (int cv = threadIdx.x; cv < cg; cv += blockDim.x) {
        const int c0 = cv * CH;
        float sc[CH], sh[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) { sc[j] = s_scale[c0 + j]; sh[j] = s_shift[c0 + j]; }
        auto act = [&](RowVec<H> z) {
#pragma unroll
            for (int j = 0; j < CH; j++) z.v[j] = fmaxf(fmaf(z.v[j], sc[j], sh[j]), 0.f);
            return z;
        };
        if (!p.pool) {
            const int stride = gridDim.x * blockDim.y;
            int m = blockIdx.x * blockDim.y + threadIdx.y;
            for (; m + (UNR - 1) * stride < p.M; m += UNR * stride) {
                RowVec<H> z[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) z[u] = ldv<H>(p.Z, z0 + RV(m + u * stride, p.M) * p.C + c0);
#pragma unroll
                for (int u = 0; u < UNR; u++) stv<H>(p.Y, y0 + RV(m + u * stride, p.M) * p.ldy + c0, act(z[u]));
            }
            for (; m < p.M; m += stride) stv<H>(p.Y, y0 + RV(m, p.M) * p.ldy + c0, act(ldv<H>(p.Z, z0 + RV(m, p.M) * p.C + c0)));
        } else {
            const long long p0 = g * p.p_gs;
            const int Hp = p.H >> 1, Wp = p.W >> 1;
            const int Mp = p.M >> 2;
            for (int mq = blockIdx.x * blockDim.y + threadIdx.y; mq < Mp; mq += gridDim.x * blockDim.y) {
                const int mp = (int)RV(mq, Mp);
                int b, yp, xp;
                pix3(mp, Hp, Wp, b, yp, xp);
                RowVec<H> z[4], mx;
                long long mi[4];
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    mi[w] = ((long long)(b * p.H + 2 * yp + (w >> 1))) * p.W + 2 * xp + (w & 1);
                    z[w] = ldv<H>(p.Z, z0 + mi[w] * p.C + c0);
                }
#pragma unroll
                for (int j = 0; j < CH; j++) mx.v[j] = 0.f;     // post-ReLU values are >= 0
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const RowVec<H> y = act(z[w]);
                    stv<H>(p.Y, y0 + mi[w] * p.ldy + c0, y);
#pragma unroll
                    for (int j = 0; j < CH; j++) mx.v[j] = fmaxf(mx.v[j], y.v[j]);
                }
                stv<H>(p.P, p0 + (long long)mp * p.C + c0, mx);      // max of the fp32 values, rounded once (rounding is monotone)
            }
        }
    }
}

// ---- BatchNorm backward through ReLU.  dzhat = dy * [z*scale+shift > 0];  xhat = (z-mean)*invstd
//      pass 1: sums[g][0][c] = sum dzhat, sums[g][1][c] = sum dzhat*xhat  (double atomics).  H: Z and dY are fp16.
//      FUSED (last unit): dY[m][c] = sum_j dout[m][j] * w_out[j][c] is formed from the staged loss gradient and the 1x1 output conv's
//      own weight / bias gradients are reduced alongside: dW_out[j][c] += dout[m][j] * relu(bn(z))[m][c]
template <bool FUSED, bool H>
__global__ void __launch_bounds__(256, FUSED ? 2 : 4) k_bn_bwd_reduce(const VvBnBwd p) {
    vv_pdl_wait();
    constexpr int CH = RowVec<H>::CH;
    constexpr int UNR = 4;
    extern __shared__ float sm[];   // [2][blockDim.y][C] partials (+ FUSED: [3][blockDim.y][C] + [blockDim.y][4])
    const int g = blockIdx.y;
    const long long z0 = g * p.z_gs;
    const long long dy0 = g * p.dy_gs + p.dy_coff;
    const float *sv = p.save + g * p.save_gs;
    const int cg = p.C / CH;
    float *ps = sm, *pq = sm + blockDim.y * p.C;
    const long long last = p.M - 1;
    auto RV = [&](int r) -> long long { return p.rev_reduce ? last - r : (long long)r; };
    for (int cv = threadIdx.x; cv < cg; cv += blockDim.x) {
        const int c0 = cv * CH;
        float sc[CH], sh[CH], mu[CH], is[CH], s[CH], q[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) {
            sc[j] = sv[c0 + j]; sh[j] = sv[p.C + c0 + j]; mu[j] = sv[2 * p.C + c0 + j]; is[j] = sv[3 * p.C + c0 + j];
            s[j] = 0.f; q[j] = 0.f;
        }
        const int stride = gridDim.x * blockDim.y;
        int m = blockIdx.x * blockDim.y + threadIdx.y;
        if (!FUSED) {
            auto acc = [&](const RowVec<H> &z, const RowVec<H> &d) {
#pragma unroll
                for (int j = 0; j < CH; j++) {
                    const float dm = fmaf(z.v[j], sc[j], sh[j]) > 0.f ? d.v[j] : 0.f;
                    s[j] += dm;
                    q[j] = fmaf(dm, z.v[j] - mu[j], q[j]);        // invstd is applied once, below
                }
            };
            for (; m + (UNR - 1) * stride < p.M; m += UNR * stride) {      // 2 * UNR independent 16-byte loads in flight per thread
                RowVec<H> z[UNR], d[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    z[u] = ldv<H>(p.Z, z0 + RV(m + u * stride) * p.C + c0);
                    d[u] = ldv<H>(p.dY, dy0 + RV(m + u * stride) * p.ldy + c0);
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) acc(z[u], d[u]);
            }
            for (; m < p.M; m += stride) acc(ldv<H>(p.Z, z0 + RV(m) * p.C + c0), ldv<H>(p.dY, dy0 + RV(m) * p.ldy + c0));
        } else {
            const float *P = p.params + p.slot.v[g] * p.slot_param_stride;
            const int oc = p.out_channels.v[g];
            float w[3][CH], a[3][CH], db[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int c = 0; c < CH; c++) { w[j][c] = j < oc ? P[p.ow_off + j * p.C + c0 + c] : 0.f; a[j][c] = 0.f; }
            const float4 *DO = reinterpret_cast<const float4 *>(p.dout) + (long long)g * p.M;
            auto one = [&](const RowVec<H> &z, const float4 &dd) {
                const float dj[3] = {dd.x, dd.y, dd.z};
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const float y = fmaf(z.v[c], sc[c], sh[c]);
                    const float u = fmaxf(y, 0.f);
                    const float d = dj[0] * w[0][c] + dj[1] * w[1][c] + dj[2] * w[2][c];
#pragma unroll
                    for (int j = 0; j < 3; j++) a[j][c] = fmaf(dj[j], u, a[j][c]);
                    const float dm = y > 0.f ? d : 0.f;
                    s[c] += dm;
                    q[c] = fmaf(dm, z.v[c] - mu[c], q[c]);
                }
#pragma unroll
                for (int j = 0; j < 3; j++) db[j] += dj[j];
            };
            for (; m + (UNR - 1) * stride < p.M; m += UNR * stride) {
                RowVec<H> z[UNR];
                float4 dd[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    z[u] = ldv<H>(p.Z, z0 + RV(m + u * stride) * p.C + c0);
                    dd[u] = DO[RV(m + u * stride)];
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) one(z[u], dd[u]);
            }
            for (; m < p.M; m += stride) one(ldv<H>(p.Z, z0 + RV(m) * p.C + c0), DO[RV(m)]);
            // block-level reduction of the 1x1 conv's gradients: [3][rows][C] + [rows][4] behind the two BN partial arrays
            float *pw = sm + 2 * blockDim.y * p.C, *pb = pw + 3 * blockDim.y * p.C;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int c = 0; c < CH; c++) pw[(j * blockDim.y + threadIdx.y) * p.C + c0 + c] = a[j][c];
            if (cv == 0) { pb[threadIdx.y * 4 + 0] = db[0]; pb[threadIdx.y * 4 + 1] = db[1]; pb[threadIdx.y * 4 + 2] = db[2]; }
        }
#pragma unroll
        for (int j = 0; j < CH; j++) { ps[threadIdx.y * p.C + c0 + j] = s[j]; pq[threadIdx.y * p.C + c0 + j] = q[j] * is[j]; }
    }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    if (FUSED) {
        const int oc = p.out_channels.v[g];
        float *Gp = p.grads + p.slot.v[g] * p.slot_param_stride;
        const float *pw = sm + 2 * blockDim.y * p.C, *pb = pw + 3 * blockDim.y * p.C;
        for (int i = tid; i < oc * p.C; i += nthr) {
            const int j = i / p.C, c = i - j * p.C;
            float a = 0.f;
            for (int r = 0; r < blockDim.y; r++) a += pw[(j * blockDim.y + r) * p.C + c];
            atomicAdd(Gp + p.ow_off + i, a);
        }
        if (tid < oc) {
            float a = 0.f;
            for (int r = 0; r < blockDim.y; r++) a += pb[r * 4 + tid];
            atomicAdd(Gp + p.ob_off + tid, a);
        }
    }
    double *sums = p.sums + g * p.sums_gs;
    for (int c = tid; c < p.C; c += nthr) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < blockDim.y; r++) { a += ps[r * p.C + c]; b += pq[r * p.C + c]; }
        atomicAdd(&sums[c], (double)a);
        atomicAdd(&sums[p.C + c], (double)b);
    }
}

//      pass 2: dz = scale * (dzhat - mean(dzhat) - xhat * mean(dzhat*xhat));  d gamma = sum dzhat*xhat, d beta = sum dzhat
//      H: Z, dY and dZ are fp16
template <bool FUSED, bool H>
__global__ void __launch_bounds__(256, FUSED ? 3 : 4) k_bn_bwd_apply(const VvBnBwd p) {
    vv_pdl_wait();
    constexpr int CH = RowVec<H>::CH;
    constexpr int UNR = 4;
    extern __shared__ float sm[];   // k1[C], k2[C]
    const int g = blockIdx.y;
    const double *sums = p.sums + g * p.sums_gs;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    float *k1 = sm, *k2 = sm + p.C;
    for (int c = tid; c < p.C; c += nthr) {
        double a = sums[c], b = sums[p.C + c];
        k1[c] = (float)(a / (double)p.M);
        k2[c] = (float)(b / (double)p.M);
        if (blockIdx.x == 0) {
            float *G = p.grads + p.slot.v[g] * p.slot_param_stride;
            G[p.gamma_off + c] = (float)b * p.grad_unscale;
            G[p.beta_off + c] = (float)a * p.grad_unscale;
        }
    }
    __syncthreads();
    const long long z0 = g * p.z_gs;
    const long long dy0 = g * p.dy_gs + p.dy_coff;
    const long long dz0 = g * p.dz_gs;
    const float *sv = p.save + g * p.save_gs;
    const int cg = p.C / CH;
    const float ss = p.store_scale;
    const long long last = p.M - 1;          // rev: rows are walked from the END (serpentine order against the previous kernel: its last rows are still in L2)
    auto RV = [&](int r) -> long long { return p.rev_apply ? last - r : (long long)r; };
    for (int cv = threadIdx.x; cv < cg; cv += blockDim.x) {
        const int c0 = cv * CH;
        // dz = scs * (dzhat - a1 - (z - mu) * is * a2) = scs * dzhat + kb * z + kc with the per-channel constants folded once
        float sc[CH], scs[CH], sh[CH], kb[CH], kc[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) {
            sc[j] = sv[c0 + j]; sh[j] = sv[p.C + c0 + j];
            const float mu = sv[2 * p.C + c0 + j], is = sv[3 * p.C + c0 + j], a1 = k1[c0 + j], a2 = k2[c0 + j];
            scs[j] = sc[j] * ss;                              // store_scale is a power of two: exact
            kb[j] = -scs[j] * is * a2;
            kc[j] = -scs[j] * a1 - kb[j] * mu;
        }
        auto dz_of = [&](const RowVec<H> &z, const RowVec<H> &d) {
            RowVec<H> o;
#pragma unroll
            for (int j = 0; j < CH; j++)
                o.v[j] = fmaf(scs[j], fmaf(z.v[j], sc[j], sh[j]) > 0.f ? d.v[j] : 0.f, fmaf(kb[j], z.v[j], kc[j]));
            return o;
        };
        const int stride = gridDim.x * blockDim.y;
        int m = blockIdx.x * blockDim.y + threadIdx.y;
        if (FUSED) {                                         // fused 1x1 output conv backward: dY formed from the staged loss gradient
            const float *P = p.params + p.slot.v[g] * p.slot_param_stride;
            const int oc = p.out_channels.v[g];
            float w[3][CH];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int c = 0; c < CH; c++) w[j][c] = j < oc ? P[p.ow_off + j * p.C + c0 + c] : 0.f;
            const float4 *DO = reinterpret_cast<const float4 *>(p.dout) + (long long)g * p.M;
            auto dy_of = [&](const float4 &dd) {
                RowVec<H> d;
#pragma unroll
                for (int c = 0; c < CH; c++) d.v[c] = dd.x * w[0][c] + dd.y * w[1][c] + dd.z * w[2][c];
                return d;
            };
            for (; m + (UNR - 1) * stride < p.M; m += UNR * stride) {
                RowVec<H> z[UNR];
                float4 dd[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    z[u] = ldv<H>(p.Z, z0 + RV(m + u * stride) * p.C + c0);
                    dd[u] = DO[RV(m + u * stride)];
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) stv<H>(p.dZ, dz0 + RV(m + u * stride) * p.C + c0, dz_of(z[u], dy_of(dd[u])));
            }
            for (; m < p.M; m += stride) stv<H>(p.dZ, dz0 + RV(m) * p.C + c0, dz_of(ldv<H>(p.Z, z0 + RV(m) * p.C + c0), dy_of(DO[RV(m)])));
        } else {
            for (; m + (UNR - 1) * stride < p.M; m += UNR * stride) {      // 2 * UNR independent 16-byte loads in flight per thread
                RowVec<H> z[UNR], d[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    z[u] = ldv<H>(p.Z, z0 + RV(m + u * stride) * p.C + c0);
                    d[u] = ldv<H>(p.dY, dy0 + RV(m + u * stride) * p.ldy + c0);
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) stv<H>(p.dZ, dz0 + RV(m + u * stride) * p.C + c0, dz_of(z[u], d[u]));   // (dZ may alias dY: own rows only)
            }
            for (; m < p.M; m += stride)
                stv<H>(p.dZ, dz0 + RV(m) * p.C + c0, dz_of(ldv<H>(p.Z, z0 + RV(m) * p.C + c0), ldv<H>(p.dY, dy0 + RV(m) * p.ldy + c0)));
        }
    }
}

// ---- MaxPool2d(2) backward: the gradient of each pooled element is added to the FIRST maximum of its 2x2 window in
//      row-major order (ATen max_pool2d keeps the first index on ties, which are frequent after ReLU).  H16: Y, dP and dY are fp16.
template <bool H16>
__global__ void k_maxpool_bwd(const void *__restrict__ Y, long long y_gs, int ldy, int y_coff, const void *__restrict__ dP,
                              long long dp_gs, void *__restrict__ dY, long long dy_gs, int lddy, int dy_coff, int B, int H, int W,
                              int C) {
    vv_pdl_wait();
    constexpr int CH = RowVec<H16>::CH;
    const int g = blockIdx.y;
    const int Hp = H >> 1, Wp = W >> 1, cg = C / CH;
    const long long total = (long long)B * Hp * Wp * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cg) * CH;
        const int mp = (int)(i / cg);
        int b, yp, xp;
        pix3(mp, Hp, Wp, b, yp, xp);
        const RowVec<H16> dp = ldv<H16>(dP, g * dp_gs + (long long)mp * C + c0);
        RowVec<H16> v[4], cur[4];
        long long mi[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            mi[w] = ((long long)(b * H + 2 * yp + (w >> 1))) * W + 2 * xp + (w & 1);
            v[w] = ldv<H16>(Y, g * y_gs + mi[w] * ldy + y_coff + c0);
            cur[w] = ldv<H16>(dY, g * dy_gs + mi[w] * lddy + dy_coff + c0);
        }
#pragma unroll
        for (int j = 0; j < CH; j++) {
            int arg = 0;
            float best = v[0].v[j];
#pragma unroll
            for (int w = 1; w < 4; w++)
                if (v[w].v[j] > best) { best = v[w].v[j]; arg = w; }
#pragma unroll
            for (int w = 0; w < 4; w++) cur[w].v[j] += (arg == w) ? dp.v[j] : 0.f;
        }
#pragma unroll
        for (int w = 0; w < 4; w++) stv<H16>(dY, g * dy_gs + mi[w] * lddy + dy_coff + c0, cur[w]);
    }
}

// ---- column sum of a (strided) gradient tensor -> ConvTranspose2d bias gradient
template <bool H>
__global__ void k_colsum(const void *__restrict__ D, long long d_gs, int ld, int coff, int M, int C, float scale, float *__restrict__ grads,
                         VvIntG slot, long long slot_stride, long long off) {
    vv_pdl_wait();
    constexpr int CH = RowVec<H>::CH;
    extern __shared__ float sm[];   // [blockDim.y][C]
    const int g = blockIdx.y;
    const int cg = C / CH;
    for (int cv = threadIdx.x; cv < cg; cv += blockDim.x) {
        const int c0 = cv * CH;
        float s[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) s[j] = 0.f;
        const int stride = gridDim.x * blockDim.y;
        int m = blockIdx.x * blockDim.y + threadIdx.y;
        for (; m + 3 * stride < M; m += 4 * stride) {
            RowVec<H> d[4];
#pragma unroll
            for (int u = 0; u < 4; u++) d[u] = ldv<H>(D, g * d_gs + (long long)(m + u * stride) * ld + coff + c0);
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int j = 0; j < CH; j++) s[j] += d[u].v[j];
        }
        for (; m < M; m += stride) {
            const RowVec<H> d = ldv<H>(D, g * d_gs + (long long)m * ld + coff + c0);
#pragma unroll
            for (int j = 0; j < CH; j++) s[j] += d.v[j];
        }
#pragma unroll
        for (int j = 0; j < CH; j++) sm[threadIdx.y * C + c0 + j] = s[j];
    }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x * blockDim.y) {
        float a = 0.f;
        for (int r = 0; r < blockDim.y; r++) a += sm[r * C + c];
        atomicAdd(grads + slot.v[g] * slot_stride + off + c, a * scale);
    }
}

// ---- 1x1 output conv (model/unet.py:63-70) with fused squared error and MSE gradient (train.py:385-392, 414-427).
//      One CTA = 256 threads = one cube (S*S pixels, one pixel row of F channels per thread and pass) of one UNet: the per-cube SSE
//      is a deterministic in-CTA reduction.  U [G][B*S*S][F];  out NCHW;  dout [G][B*S*S][4].
template <bool H>
__global__ void __launch_bounds__(256) k_outconv_fwd(const VvOutFwd p) {
    vv_pdl_wait();
    extern __shared__ float sm[];   // w [4][F], b[4]
    const int g = blockIdx.y, b = blockIdx.x;
    const int F = p.F, SS = p.S * p.S;
    float *w = sm;
    float *bs = w + 4 * F;
    __shared__ float red[8];
    const int oc = p.out_channels.v[g];
    const float *P = p.params + p.slot.v[g] * p.slot_param_stride;
    for (int i = threadIdx.x; i < 4 * F; i += 256) w[i] = (i < oc * F) ? P[p.w_off + i] : 0.f;
    if (threadIdx.x < 4) bs[threadIdx.x] = threadIdx.x < oc ? P[p.b_off + threadIdx.x] : 0.f;
    __syncthreads();
    const long long u0 = g * p.u_gs + (long long)b * SS * F;
    const bool flow = p.target_is_flow.v[g] != 0;
    float *out = flow ? p.of_out : p.raw_out;
    const int out_ctot = flow ? p.of_out_channels : p.raw_out_channels;
    const int out_c0 = p.out_slot.v[g] * (flow ? 2 : 3);
    const float *tgt = nullptr;
    if (p.sse) {
        tgt = flow ? p.x_of + ((long long)b * p.x_of_channels + 2 * p.target_index.v[g]) * SS
                   : p.x + ((long long)b * p.x_channels + 3 * p.target_index.v[g]) * SS;
    }
    const float coef = flow ? p.coef_of : p.coef_raw;
    float sse = 0.f;
    // one thread per pixel: its F channels are one contiguous NHWC row read with 8- / 16-byte loads (a warp reads 32 consecutive
    // rows), the three dot products run against the weights broadcast from shared memory in channel order (as before: same sums)
    for (int pix = threadIdx.x; pix < SS; pix += 256) {
        float o[4] = {bs[0], bs[1], bs[2], bs[3]};
        const long long r0 = u0 + (long long)pix * F;
        for (int c = 0; c < F; c += 8) {
            float uu[8];
            ld8<H>(p.U, r0 + c, uu);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                o[0] = fmaf(uu[k], w[c + k], o[0]); o[1] = fmaf(uu[k], w[F + c + k], o[1]); o[2] = fmaf(uu[k], w[2 * F + c + k], o[2]);
            }
        }
        float dv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < oc; j++) {
            if (out) out[((long long)b * out_ctot + out_c0 + j) * SS + pix] = o[j];
            if (tgt) {
                float e = o[j] - tgt[(long long)j * SS + pix];
                sse += e * e;
                dv[j] = coef * e;
            }
        }
        if (p.dout)
            *reinterpret_cast<float4 *>(p.dout + ((long long)g * p.B * SS + (long long)b * SS + pix) * 4) =
                make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
    if (p.sse) {
        for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sse;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int i = 0; i < 8; i++) s += red[i];
            p.sse[g * p.B + b] = s;
        }
    }
}

// backward of the 1x1 conv: dU[m][c] = sum_j dout[m][j] w[j][c];  dW[j][c] += sum_m dout[m][j] U[m][c];  db[j] += sum_m dout[m][j]
// thread = (pixel slot, 4 channels): float4 loads of U / stores of dU, 256/(F/4) pixels per block iteration, two iterations in
// flight; the weight / bias gradients are reduced in shared memory and flushed with one atomic per element per block.
template <bool H>
__global__ void __launch_bounds__(256) k_outconv_bwd(const VvOutBwd p) {
    vv_pdl_wait();
    extern __shared__ float sm_ob[];            // dW partial [3][F], db partial [4]
    const int g = blockIdx.y;
    const int F = p.F, FQ = F >> 2;
    const int q = threadIdx.x % FQ, ps = threadIdx.x / FQ, npix = 256 / FQ;
    const int oc = p.out_channels.v[g];
    const float *P = p.params + p.slot.v[g] * p.slot_param_stride;
    float *G = p.grads + p.slot.v[g] * p.slot_param_stride;
    const long long u0 = g * p.u_gs;
    const long long du0 = g * p.du_gs;
    const float dus = p.du_scale;
    const int SS = p.S * p.S;
    const bool flow = p.target_is_flow.v[g] != 0;
    const float *ext = flow ? p.grad_of_out : p.grad_raw_out;      // external NCHW gradient, else the staged [m][4]
    const int ext_ctot = flow ? p.of_out_channels : p.raw_out_channels;
    const int ext_c0 = p.out_slot.v[g] * (flow ? 2 : 3);
    for (int i = threadIdx.x; i < 3 * F + 4; i += 256) sm_ob[i] = 0.f;
    __syncthreads();
    float4 w[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
        w[j] = j < oc ? *reinterpret_cast<const float4 *>(P + p.w_off + j * F + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a[3];
#pragma unroll
    for (int j = 0; j < 3; j++) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float db[3] = {0.f, 0.f, 0.f};
    auto load_d = [&](int m, float *d) {
        if (ext) {
            int b = m / SS, pix = m - b * SS;
            const float *e = ext + ((long long)b * ext_ctot + ext_c0) * SS + pix;
            d[0] = e[0];
            d[1] = oc > 1 ? e[SS] : 0.f;
            d[2] = oc > 2 ? e[2 * SS] : 0.f;
        } else {
            float4 v = *reinterpret_cast<const float4 *>(p.dout + ((long long)g * p.M + m) * 4);
            d[0] = v.x; d[1] = v.y; d[2] = v.z;
        }
    };
    auto body = [&](int m, const float *d, const float4 &u) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            a[j].x = fmaf(d[j], u.x, a[j].x); a[j].y = fmaf(d[j], u.y, a[j].y);
            a[j].z = fmaf(d[j], u.z, a[j].z); a[j].w = fmaf(d[j], u.w, a[j].w);
            db[j] += d[j];
        }
        float4 o;
        o.x = (d[0] * w[0].x + d[1] * w[1].x + d[2] * w[2].x) * dus; o.y = (d[0] * w[0].y + d[1] * w[1].y + d[2] * w[2].y) * dus;
        o.z = (d[0] * w[0].z + d[1] * w[1].z + d[2] * w[2].z) * dus; o.w = (d[0] * w[0].w + d[1] * w[1].w + d[2] * w[2].w) * dus;
        st4<H>(p.dU, du0 + (long long)m * F + 4 * q, o);
    };
    const int stride = gridDim.x * npix;
    int m = blockIdx.x * npix + ps;
    for (; m + stride < p.M; m += 2 * stride) {          // two pixels in flight per thread
        float d0[3], d1[3];
        load_d(m, d0);
        load_d(m + stride, d1);
        const float4 ua = ld4<H>(p.U, u0 + (long long)m * F + 4 * q);
        const float4 ub = ld4<H>(p.U, u0 + (long long)(m + stride) * F + 4 * q);
        body(m, d0, ua);
        body(m + stride, d1, ub);
    }
    if (m < p.M) {
        float d0[3];
        load_d(m, d0);
        const float4 ua = ld4<H>(p.U, u0 + (long long)m * F + 4 * q);
        body(m, d0, ua);
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        atomicAdd(&sm_ob[j * F + 4 * q + 0], a[j].x); atomicAdd(&sm_ob[j * F + 4 * q + 1], a[j].y);
        atomicAdd(&sm_ob[j * F + 4 * q + 2], a[j].z); atomicAdd(&sm_ob[j * F + 4 * q + 3], a[j].w);
        if (q == 0) atomicAdd(&sm_ob[3 * F + j], db[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * F; i += 256)
        if (i / F < oc) atomicAdd(G + p.w_off + i, sm_ob[i]);
    if (threadIdx.x < oc) atomicAdd(G + p.b_off + threadIdx.x, sm_ob[3 * F + threadIdx.x]);
}

// ---- gradient re-layout back to PyTorch's parameter layouts
__global__ void k_scatter_conv_wgrad(const float *__restrict__ dWf, long long gs, int N, int C, int Cp, float *__restrict__ grads,
                                     VvIntG slot, long long slot_stride, long long w_off) {
    vv_pdl_wait();
    const int g = blockIdx.y;
    const int total = N * C * 9;
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // i indexes the PyTorch layout [n][c][t] (coalesced writes)
    if (i >= total) return;
    int t = i % 9, c = (i / 9) % C, n = i / (9 * C);
    grads[slot.v[g] * slot_stride + w_off + i] = dWf[g * gs + ((long long)t * N + n) * Cp + c];
}

// tiled variant: block = 32 output channels x 32 input channels x 9 taps through shared memory, coalesced on both sides
__global__ void __launch_bounds__(256) k_scatter_conv_wgrad_tiled(const float *__restrict__ dWf, long long gs, int N, int C, int Cp,
                                                                  float *__restrict__ grads, VvIntG slot, long long slot_stride,
                                                                  long long w_off) {
    vv_pdl_wait();
    __shared__ float tile[32][32 * 9 + 1];
    const int g = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int cw = min(32, C - c0);
    if (cw <= 0) return;
    for (int i = threadIdx.x; i < 9 * 32 * 32; i += 256) {
        const int c = i & 31, n = (i >> 5) & 31, t = i >> 10;
        tile[n][c * 9 + t] = dWf[g * gs + ((long long)t * N + n0 + n) * Cp + c0 + c];
    }
    __syncthreads();
    float *G = grads + slot.v[g] * slot_stride + w_off;
    for (int i = threadIdx.x; i < 32 * 288; i += 256) {
        const int n = i / 288, r = i - n * 288;
        if (r < cw * 9) G[((long long)(n0 + n) * C + c0) * 9 + r] = tile[n][r];
    }
}

__global__ void k_scatter_ct_wgrad(const float *__restrict__ dWb, long long gs, int Ci, int Co, float scale, float *__restrict__ grads,
                                   VvIntG slot, long long slot_stride, long long w_off) {
    vv_pdl_wait();
    const int g = blockIdx.y;
    const int total = Ci * Co * 9;
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // PyTorch layout [ci][co][ky][kx]
    if (i >= total) return;
    int kx = i % 3, ky = (i / 3) % 3, co = (i / 9) % Co, ci = i / (9 * Co);
    // ky = py + 1 - 2 sy  ->  ky=0:(py=1,sy=1)  ky=1:(py=0,sy=0)  ky=2:(py=1,sy=0)
    int py = (ky == 1) ? 0 : 1, sy = (ky == 0) ? 1 : 0;
    int px = (kx == 1) ? 0 : 1, sx = (kx == 0) ? 1 : 0;
    int s = sy * 2 + sx, ph = py * 2 + px;
    grads[slot.v[g] * slot_stride + w_off + i] = dWb[g * gs + ((long long)s * 4 * Co + ph * Co + co) * Ci + ci] * scale;
}

// ---- losses from the per-cube SSE buffer: mean over B * ch * S * S (train.py:385-392)
__global__ void k_losses(const float *__restrict__ sse, int G, int B, VvIntG is_flow, float inv_raw, float inv_of, float *__restrict__ out) {
    vv_pdl_wait();
    __shared__ double r[2][256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < G * B; i += 256) {
        int g = i / B;
        if (is_flow.v[g]) b += sse[i]; else a += sse[i];
    }
    r[0][threadIdx.x] = a; r[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { r[0][threadIdx.x] += r[0][threadIdx.x + s]; r[1][threadIdx.x] += r[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = (float)(r[0][0] * inv_raw); out[1] = (float)(r[1][0] * inv_of); }
}

// ---- Adam (torch.optim.Adam, amsgrad=False, maximize=False): train.py:376
// blockIdx.y = slot: the same [0, n) range of every slot, slot_stride floats apart (0: one flat range)
__global__ void k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, long long n,
                       long long slot_stride, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
    vv_pdl_wait();
    { const long long o = blockIdx.y * slot_stride; p += o; g += o; m += o; v += o; }
    long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 pp = *reinterpret_cast<float4 *>(p + i);
        float4 gg = *reinterpret_cast<const float4 *>(g + i);
        float4 mm = *reinterpret_cast<float4 *>(m + i);
        float4 vv = *reinterpret_cast<float4 *>(v + i);
        float *pa = &pp.x, *ga = &gg.x, *ma = &mm.x, *va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float gr = ga[j] * gscale + wd * pa[j];
            ma[j] = b1 * ma[j] + (1.f - b1) * gr;
            va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
            float denom = sqrtf(va[j]) / bc2_sqrt + eps;
            pa[j] -= (lr / bc1) * (ma[j] / denom);
        }
        *reinterpret_cast<float4 *>(p + i) = pp;
        *reinterpret_cast<float4 *>(m + i) = mm;
        *reinterpret_cast<float4 *>(v + i) = vv;
    } else {
        for (; i < n; i++) {
            float gr = g[i] * gscale + wd * p[i];
            m[i] = b1 * m[i] + (1.f - b1) * gr;
            v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
            float denom = sqrtf(v[i]) / bc2_sqrt + eps;
            p[i] -= (lr / bc1) * (m[i] / denom);
        }
    }
}

// ---- cube staging: uint8 [N,T,S,S,3] -> x [N,3T,S,S] (/255), flow [N,To,S,S,2] -> x_of [N,2To,S,S]   (vad_datasets.py:153-165)
__global__ void k_cubes_to_tensors(const uint8_t *__restrict__ raw, const float *__restrict__ flow, float *__restrict__ x,
                                   float *__restrict__ x_of, int n, int T, int To, int S) {
    vv_pdl_wait();
    const long long SS = (long long)S * S;
    const long long nraw = (long long)n * T * 3 * SS, nof = flow ? (long long)n * To * 2 * SS : 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nraw + nof; i += (long long)gridDim.x * blockDim.x) {
        if (i < nraw) {   // i indexes the output [n][t][c][pix]
            long long pix = i % SS;
            int c = (int)((i / SS) % 3), t = (int)((i / (3 * SS)) % T);
            long long b = i / (3 * SS * T);
            x[i] = (float)raw[((b * T + t) * SS + pix) * 3 + c] / 255.0f;   // torchvision ToTensor: byte -> float32 .div(255)
        } else {
            long long k = i - nraw;
            long long pix = k % SS;
            int c = (int)((k / SS) % 2), t = (int)((k / (2 * SS)) % To);
            long long b = k / (2 * SS * To);
            x_of[k] = flow[((b * To + t) * SS + pix) * 2 + c];
        }
    }
}

static inline dim3 row_block(int C, int ch, int &rows) {      // ch channels per thread
    int tx = C / ch;
    if (tx > 64) tx = 64;
    if (tx < 1) tx = 1;
    rows = 256 / tx;
    return dim3(tx, rows);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int vv_prep_input(const float *x, void *X0, int x0_f16, int G, int B, int T, int S, int cinp, int padding, const VvIntG &erase, cudaStream_t st) {
    int M = B * S * S;
    if (x0_f16) vv_launch(k_prep_input<true>, dim3(vv_cdiv(M, 256), G), dim3(256), 0, st, x, X0, B, T, S, cinp, padding, erase);
    else vv_launch(k_prep_input<false>, dim3(vv_cdiv(M, 256), G), dim3(256), 0, st, x, X0, B, T, S, cinp, padding, erase);
    VV_CKL();
    return 0;
}

int vv_prep_conv_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, long long g_off,
                   long long beta_off, int N, int C, int Cp, void *Wf, long long wf_gs, void *Wd, long long wd_gs, int w_f16, float *vec,
                   long long vec_gs, int G, cudaStream_t st) {
    if (N % 32 == 0 && Cp % 32 == 0)
        vv_launch(k_prep_conv_w_tiled, dim3(N / 32, Cp / 32, G), dim3(256), 0, st, params, slot, slot_stride, w_off, b_off, g_off, beta_off, N, C, Cp, Wf,
                                                                    wf_gs, Wd, wd_gs, w_f16, vec, vec_gs);
    else
        vv_launch(k_prep_conv_w, dim3(vv_cdiv(9LL * N * Cp, 256), G), dim3(256), 0, st, params, slot, slot_stride, w_off, b_off, g_off, beta_off, N, C,
                                                                         Cp, Wf, wf_gs, Wd, wd_gs, w_f16, vec, vec_gs);
    VV_CKL();
    return 0;
}

static void layout_blocks(VvPrepAll &all) {
    int blk = 0;
    for (int i = 0; i < all.n; i++) {
        all.u[i].blk0 = blk;
        blk += (all.u[i].N / 32) * (all.u[i].Cp / 32);
    }
    all.total_blocks = blk;
}
int vv_prep_conv_w_all(const float *params, const VvIntG &slot, long long slot_stride, VvPrepAll &all, int G, cudaStream_t st) {
    layout_blocks(all);
    vv_launch(k_prep_conv_w_all, dim3(all.total_blocks, 1, G), dim3(256), 0, st, params, slot, slot_stride, all);
    VV_CKL();
    return 0;
}
int vv_scatter_conv_wgrad_all(float *grads, const VvIntG &slot, long long slot_stride, VvPrepAll &all, int G, cudaStream_t st) {
    layout_blocks(all);
    vv_launch(k_scatter_conv_wgrad_all, dim3(all.total_blocks, 1, G), dim3(256), 0, st, grads, slot, slot_stride, all);
    VV_CKL();
    return 0;
}

// the same re-layout through a shared-memory tile of 32 input x 32 output channels: coalesced 1152-byte reads of the ConvTranspose2d
// weight [Ci][Co][3][3], 8 / 16-byte writes of both operand layouts (the element-per-thread kernel above reads with a stride of
// 9 Co floats and writes 2-byte elements 4 Co apart: 41 us for the 256 -> 128 layer; this one 3x faster)
__global__ void __launch_bounds__(256) k_prep_ct_w_tiled(const float *__restrict__ params, VvIntG slot, long long slot_stride, long long w_off,
                                                         long long b_off, int Ci, int Co, void *__restrict__ Wbf, long long wf_gs,
                                                         void *__restrict__ Wbd, long long wd_gs, int w_f16, float *__restrict__ vec,
                                                         long long vec_gs) {
    vv_pdl_wait();
    __shared__ float tile[32][32 * 9 + 1];
    const int g = blockIdx.y;
    const int nco = Co / 32, ci0 = (blockIdx.x / nco) * 32, co0 = (blockIdx.x % nco) * 32;
    const float *P = params + slot.v[g] * slot_stride;
    for (int i = threadIdx.x; i < 32 * 288; i += 256) {
        const int ci = i / 288, r = i - ci * 288;
        tile[ci][r] = P[w_off + ((long long)(ci0 + ci) * Co + co0) * 9 + r];
    }
    __syncthreads();
    // tap s = (sy, sx) of the 2x2 gather, output phase ph = (py, px): kernel element (py + 1 - 2 sy, px + 1 - 2 sx), absent when negative
    for (int i = threadIdx.x; i < 16 * 32 * 8; i += 256) {
        const int q = i & 7, l = (i >> 3) & 31, sp = i >> 8, ph = sp & 3, t = sp >> 2;
        const int ky = (ph >> 1) + 1 - 2 * (t >> 1), kx = (ph & 1) + 1 - 2 * (t & 1);
        const bool ok = ky >= 0 && kx >= 0;
        const int kk = ok ? ky * 3 + kx : 0;
        {   // forward operand [s][ph*Co + co][ci]: l = output channel, four consecutive input channels
            const float4 v = ok ? make_float4(tile[4 * q][l * 9 + kk], tile[4 * q + 1][l * 9 + kk], tile[4 * q + 2][l * 9 + kk], tile[4 * q + 3][l * 9 + kk])
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            const long long o = g * wf_gs + ((long long)t * 4 * Co + ph * Co + co0 + l) * Ci + ci0 + 4 * q;
            if (w_f16) st4<true>(Wbf, o, v); else st4<false>(Wbf, o, v);
        }
        {   // input-gradient operand [s][ci][ph*Co + co]: l = input channel, four consecutive output channels
            const float4 v = ok ? make_float4(tile[l][(4 * q) * 9 + kk], tile[l][(4 * q + 1) * 9 + kk], tile[l][(4 * q + 2) * 9 + kk], tile[l][(4 * q + 3) * 9 + kk])
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            const long long o = g * wd_gs + ((long long)t * Ci + ci0 + l) * (4 * Co) + ph * Co + co0 + 4 * q;
            if (w_f16) st4<true>(Wbd, o, v); else st4<false>(Wbd, o, v);
        }
    }
    if (blockIdx.x == 0)
        for (int n = threadIdx.x; n < Co; n += blockDim.x) vec[g * vec_gs + n] = P[b_off + n];
}

int vv_prep_ct_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, int Ci, int Co,
                 void *Wbf, long long wf_gs, void *Wbd, long long wd_gs, int w_f16, float *vec, long long vec_gs, int G, cudaStream_t st) {
    if (Ci % 32 == 0 && Co % 32 == 0) {
        vv_launch(k_prep_ct_w_tiled, dim3((Ci / 32) * (Co / 32), G), dim3(256), 0, st, params, slot, slot_stride, w_off, b_off, Ci, Co, Wbf, wf_gs,
                  Wbd, wd_gs, w_f16, vec, vec_gs);
        VV_CKL();
        return 0;
    }
    vv_launch(k_prep_ct_w, dim3(vv_cdiv(16LL * Co * Ci, 256), G), dim3(256), 0, st, params, slot, slot_stride, w_off, b_off, Ci, Co, Wbf, wf_gs, Wbd,
                                                                     wd_gs, w_f16, vec, vec_gs);
    VV_CKL();
    return 0;
}

int vv_bn_apply(const VvBnApply &p, int G, cudaStream_t st) {
    int rows;
    VV_REQUIRE(p.C % 4 == 0, "bn_apply: channel count %d", p.C);
    dim3 blk = row_block(p.C, 4, rows);
    // ONE resident wave (148 SMs x 4 blocks) over all groups: every block pays a prologue (the per-channel BatchNorm constants, in
    // double precision) and must amortise it over many rows -- with one block per 4 rows of work these passes ran at half of the HBM rate
    int work = p.pool ? p.M / 4 : p.M;
    int gx = vv_cdiv(work, rows * 4);
    const int cap = (148 * 4) / G > 0 ? (148 * 4) / G : 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    if (p.y_f16) vv_launch(k_bn_apply<true>, dim3(gx, G), dim3(blk), 2 * p.C * sizeof(float), st, p);
    else vv_launch(k_bn_apply<false>, dim3(gx, G), dim3(blk), 2 * p.C * sizeof(float), st, p);
    VV_CKL();
    return 0;
}

int vv_bn_bwd(const VvBnBwd &p, int G, cudaStream_t st) {
    int rows;
    VV_REQUIRE(p.C % 4 == 0, "bn_bwd: channel count %d", p.C);
    dim3 blk = row_block(p.C, 4, rows);
    const int cap = (148 * 4) / G > 0 ? (148 * 4) / G : 1;      // one resident wave over all groups (see vv_bn_apply)
    int gx = vv_cdiv(p.M, rows * 8);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    if (p.dout) VV_REQUIRE(p.C % 4 == 0 && p.params && p.grads, "bn_bwd: fused output-conv backward needs params / grads");
    VV_REQUIRE(!p.dz_f16 || (const void *)p.dZ != (const void *)p.dY, "bn_bwd: an fp16 dZ cannot alias dY");
    VV_REQUIRE(p.z_f16 == p.dz_f16, "bn_bwd: Z and dZ are both fp16 (fp16 mode) or both fp32");
    if (p.dout) {
        const size_t smr = (5 * rows * p.C + 4 * rows) * sizeof(float);
        if (p.z_f16) vv_launch(k_bn_bwd_reduce<true, true>, dim3(gx, G), dim3(blk), smr, st, p);
        else vv_launch(k_bn_bwd_reduce<true, false>, dim3(gx, G), dim3(blk), smr, st, p);
    } else {
        const size_t smr = 2 * rows * p.C * sizeof(float);
        if (p.z_f16) vv_launch(k_bn_bwd_reduce<false, true>, dim3(gx, G), dim3(blk), smr, st, p);
        else vv_launch(k_bn_bwd_reduce<false, false>, dim3(gx, G), dim3(blk), smr, st, p);
    }
    VV_CKL();
    int gx2 = vv_cdiv(p.M, rows * 4);
    if (gx2 > cap) gx2 = cap;
    if (gx2 < 1) gx2 = 1;
    const dim3 grid(gx2, G);
    const size_t sm = 2 * p.C * sizeof(float);
    if (p.dout) {
        if (p.dz_f16) vv_launch(k_bn_bwd_apply<true, true>, dim3(grid), dim3(blk), sm, st, p);
        else vv_launch(k_bn_bwd_apply<true, false>, dim3(grid), dim3(blk), sm, st, p);
    } else {
        if (p.dz_f16) vv_launch(k_bn_bwd_apply<false, true>, dim3(grid), dim3(blk), sm, st, p);
        else vv_launch(k_bn_bwd_apply<false, false>, dim3(grid), dim3(blk), sm, st, p);
    }
    VV_CKL();
    return 0;
}

int vv_maxpool_bwd(const void *Y, int y_f16, long long y_gs, int ldy, int y_coff, const void *dP, long long dp_gs, void *dY, long long dy_gs,
                   int lddy, int dy_coff, int G, int B, int H, int W, int C, cudaStream_t st) {
    long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
    int gx = vv_cdiv(total, 256);
    if (gx > 148 * 16) gx = 148 * 16;
    if (y_f16) vv_launch(k_maxpool_bwd<true>, dim3(gx, G), dim3(256), 0, st, Y, y_gs, ldy, y_coff, dP, dp_gs, dY, dy_gs, lddy, dy_coff, B, H, W, C);
    else vv_launch(k_maxpool_bwd<false>, dim3(gx, G), dim3(256), 0, st, Y, y_gs, ldy, y_coff, dP, dp_gs, dY, dy_gs, lddy, dy_coff, B, H, W, C);
    VV_CKL();
    return 0;
}

int vv_colsum(const void *D, int d_f16, long long d_gs, int ld, int coff, int M, int C, float scale, float *grads, const VvIntG &slot,
              long long slot_stride, long long off, int G, cudaStream_t st) {
    int rows;
    dim3 blk = row_block(C, 4, rows);
    int gx = vv_cdiv(M, rows * 16);
    if (gx > 148 * 2) gx = 148 * 2;
    if (gx < 1) gx = 1;
    if (d_f16) vv_launch(k_colsum<true>, dim3(gx, G), dim3(blk), rows * C * sizeof(float), st, D, d_gs, ld, coff, M, C, scale, grads, slot, slot_stride, off);
    else vv_launch(k_colsum<false>, dim3(gx, G), dim3(blk), rows * C * sizeof(float), st, D, d_gs, ld, coff, M, C, scale, grads, slot, slot_stride, off);
    VV_CKL();
    return 0;
}

int vv_outconv_fwd(const VvOutFwd &p, int G, cudaStream_t st) {
    VV_REQUIRE((p.S * p.S) % 256 == 0, "outconv: S*S must be a multiple of 256 (S=%d)", p.S);
    VV_REQUIRE(p.F % 8 == 0, "outconv: feature count %d must be a multiple of 8", p.F);
    const size_t smem = (4 * p.F + 4) * sizeof(float);
    if (p.u_f16) vv_launch(k_outconv_fwd<true>, dim3(p.B, G), dim3(256), smem, st, p);
    else vv_launch(k_outconv_fwd<false>, dim3(p.B, G), dim3(256), smem, st, p);
    VV_CKL();
    return 0;
}

int vv_outconv_bwd(const VvOutBwd &p, int G, cudaStream_t st) {
    VV_REQUIRE(p.F % 16 == 0 && p.F <= 1024 && 256 % (p.F / 4) == 0, "outconv backward: unsupported features_root %d", p.F);
    const int npix = 256 / (p.F / 4);
    int gx = vv_cdiv(p.M, npix * 8);
    if (gx > 148 * 8) gx = 148 * 8;
    if (gx < 1) gx = 1;
    if (p.u_f16) vv_launch(k_outconv_bwd<true>, dim3(gx, G), dim3(256), (3 * p.F + 4) * sizeof(float), st, p);
    else vv_launch(k_outconv_bwd<false>, dim3(gx, G), dim3(256), (3 * p.F + 4) * sizeof(float), st, p);
    VV_CKL();
    return 0;
}

int vv_scatter_conv_wgrad(const float *dWf, long long gs, int N, int C, int Cp, float *grads, const VvIntG &slot, long long slot_stride,
                          long long w_off, int G, cudaStream_t st) {
    if (N % 32 == 0 && Cp % 32 == 0)
        vv_launch(k_scatter_conv_wgrad_tiled, dim3(N / 32, Cp / 32, G), dim3(256), 0, st, dWf, gs, N, C, Cp, grads, slot, slot_stride, w_off);
    else
        vv_launch(k_scatter_conv_wgrad, dim3(vv_cdiv(9LL * N * C, 256), G), dim3(256), 0, st, dWf, gs, N, C, Cp, grads, slot, slot_stride, w_off);
    VV_CKL();
    return 0;
}

int vv_scatter_ct_wgrad(const float *dWb, long long gs, int Ci, int Co, float scale, float *grads, const VvIntG &slot, long long slot_stride,
                        long long w_off, int G, cudaStream_t st) {
    vv_launch(k_scatter_ct_wgrad, dim3(vv_cdiv(9LL * Ci * Co, 256), G), dim3(256), 0, st, dWb, gs, Ci, Co, scale, grads, slot, slot_stride, w_off);
    VV_CKL();
    return 0;
}

int vv_losses(const float *sse, int G, int B, const VvIntG &is_flow, float inv_raw, float inv_of, float *out, cudaStream_t st) {
    vv_launch(k_losses, dim3(1), dim3(256), 0, st, sse, G, B, is_flow, inv_raw, inv_of, out);
    VV_CKL();
    return 0;
}

extern "C" int vecvad_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, float grad_scale, vecvad_stream stream) {
    VV_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam: bad arguments");
    VV_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
               "adam: buffers must be 16-byte aligned");
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    long long nthr = (n + 3) / 4;
    vv_launch(k_adam, dim3(vv_cdiv(nthr, 256)), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, n, 0LL, lr, beta1, beta2, eps,
              weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
    VV_CKL();
    return 0;
}

// the same update on [begin, end) of every one of n_slots slots (the gradient phases of vecvad_net_grad_phase_ranges)
extern "C" int vecvad_adam_step_ranges(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int n_slots, int64_t slot_stride,
                                       int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                       float grad_scale, vecvad_stream stream) {
    VV_REQUIRE(params && grads && exp_avg && exp_avg_sq && n_slots >= 1 && step >= 1, "adam_ranges: bad arguments");
    VV_REQUIRE(0 <= begin && begin < end && end <= slot_stride && begin % 4 == 0 && slot_stride % 4 == 0, "adam_ranges: bad range [%lld, %lld) of %lld",
               (long long)begin, (long long)end, (long long)slot_stride);
    VV_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "adam_ranges: buffers must be 16-byte aligned");
    VV_REQUIRE(n_slots <= 65535, "adam_ranges: too many slots");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const long long n = end - begin, nthr = (n + 3) / 4;
    vv_launch(k_adam, dim3(vv_cdiv(nthr, 256), n_slots), dim3(256), 0, (cudaStream_t)stream, params + begin, grads + begin, exp_avg + begin,
              exp_avg_sq + begin, (long long)n, (long long)slot_stride, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
    VV_CKL();
    return 0;
}

extern "C" int vecvad_cubes_to_tensors(const uint8_t *raw, const float *flow, float *x, float *x_of, int n, int t_raw, int t_of, int patch,
                                       vecvad_stream stream) {
    VV_REQUIRE(raw && x && n > 0 && t_raw > 0 && patch > 0, "cubes_to_tensors: bad arguments");
    VV_REQUIRE(!flow || x_of, "cubes_to_tensors: x_of missing");
    long long total = (long long)n * patch * patch * (3 * t_raw + (flow ? 2 * t_of : 0));
    int gx = vv_cdiv(total, 256);
    if (gx > 148 * 16) gx = 148 * 16;
    vv_launch(k_cubes_to_tensors, dim3(gx), dim3(256), 0, (cudaStream_t)stream, raw, flow, x, x_of, n, t_raw, t_of, patch);
    VV_CKL();
    return 0;
}

namespace {
__global__ void k_f32_to_f16(const float *__restrict__ src, int ld, int cols, long long rows, __half *__restrict__ dst) {
    vv_pdl_wait();
    const int cq = cols >> 2;
    const long long total = rows * cq;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cq;
        const int c4 = (int)(i - r * cq);
        const float4 v = *reinterpret_cast<const float4 *>(src + r * ld + c4 * 4);
        *reinterpret_cast<uint2 *>(dst + r * cols + c4 * 4) = pack_h4(v);
    }
}
}  // namespace

namespace {
__global__ void k_f16_to_f32(const __half *__restrict__ src, long long n, float *__restrict__ dst) {
    vv_pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = __half2float(src[i]);
}
}  // namespace

int vv_f16_to_f32(const void *src, long long n, float *dst, cudaStream_t st) {
    int gx = vv_cdiv(n, 256);
    if (gx > 148 * 16) gx = 148 * 16;
    if (gx < 1) gx = 1;
    vv_launch(k_f16_to_f32, dim3(gx), dim3(256), 0, st, (const __half *)src, n, dst);
    VV_CKL();
    return 0;
}

int vv_f32_to_f16(const float *src, int ld, int cols, long long rows, void *dst, cudaStream_t st) {
    VV_REQUIRE(cols % 4 == 0 && ld % 4 == 0, "f32_to_f16: cols / ld must be multiples of 4");
    long long total = rows * (cols / 4);
    int gx = vv_cdiv(total, 256);
    if (gx > 148 * 16) gx = 148 * 16;
    if (gx < 1) gx = 1;
    vv_launch(k_f32_to_f16, dim3(gx), dim3(256), 0, st, src, ld, cols, rows, (__half *)dst);
    VV_CKL();
    return 0;
}

