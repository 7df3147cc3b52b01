// Bandwidth-bound kernels of the completion-UNet set: input staging, weight re-layout, BatchNorm
// (train/eval) + ReLU (+2x2 max-pool) apply, BatchNorm backward, max-pool backward, 1x1 output conv
// with fused squared-error / MSE gradient, Adam.  All operate on grouped NHWC fp32 tensors
// [G][B*H*W][ld] (blockIdx.z / .y = UNet index g).
//
// Reference semantics restated (reference file:line in each kernel's comment); PyTorch layer
// definitions at model/unet.py:9-16 (conv+BN+ReLU), :38 (MaxPool2d(2)), :54 (ConvTranspose2d), :66 (1x1 conv).
#include "unet_kernels.h"

namespace {

__device__ __forceinline__ void pix3(int m, int H, int W, int &b, int &y, int &x) {
    x = m % W;
    int t = m / W;
    y = t % H;
    b = t / H;
}

// ---- input staging: x NCHW [B, 3*T, S, S] -> X0 [G][B*S*S][cinp] with frame erase[g] dropped or zeroed (model/unet.py:179-183)
__global__ void k_prep_input(const float *__restrict__ x, float *__restrict__ X0, int B, int T, int S, int cinp, int padding,
                             VvIntG erase) {
    const int g = blockIdx.y;
    const int M = B * S * S;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int pix = m % (S * S), b = m / (S * S);
    const int e = erase.v[g];
    const float *xb = x + (long long)b * 3 * T * S * S + pix;
    float *dst = X0 + ((long long)g * M + m) * cinp;
    const int creal = padding ? 3 * T : 3 * (T - 1);
    for (int c4 = 0; c4 < cinp; c4 += 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int c = c4 + j;
            float val = 0.f;
            if (c < creal) {
                int src = padding ? c : (c < 3 * e ? c : c + 3);
                bool erased = padding && (c >= 3 * e) && (c < 3 * e + 3);
                if (!erased) val = __ldg(xb + (long long)src * S * S);
            }
            v[j] = val;
        }
        *reinterpret_cast<float4 *>(dst + c4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ---- weight re-layout, once per step.  Conv2d weight [N][C][3][3] -> Wf[t][N][Cp] (forward B operand, K contiguous)
//      and Wd[t][C][N] (dgrad B operand); bias / gamma / beta gathered into a uniform-stride vector block [3][N].
__global__ void k_prep_conv_w(const float *__restrict__ params, VvIntG slot, long long slot_stride, long long w_off, long long b_off,
                              long long g_off, long long beta_off, int N, int C, int Cp, float *__restrict__ Wf, long long wf_gs,
                              float *__restrict__ Wd, long long wd_gs, float *__restrict__ vec, long long vec_gs) {
    const int g = blockIdx.y;
    const float *P = params + slot.v[g] * slot_stride;
    const int total = 9 * N * Cp;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        int c = i % Cp;
        int n = (i / Cp) % N;
        int t = i / (Cp * N);
        float v = (c < C) ? P[w_off + ((long long)n * C + c) * 9 + t] : 0.f;
        Wf[g * wf_gs + i] = v;
        if (Wd && c < C) Wd[g * wd_gs + ((long long)t * C + c) * N + n] = v;
    }
    if (blockIdx.x == 0 && vec) {
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            vec[g * vec_gs + n] = P[b_off + n];
            vec[g * vec_gs + N + n] = P[g_off + n];
            vec[g * vec_gs + 2 * N + n] = P[beta_off + n];
        }
    }
}

// ConvTranspose2d(k3,s2,p1,op1) weight [Ci][Co][3][3] -> 2x2-tap "big" matrices over the 4 output phases:
//   out[2y+py, 2x+px, co] = sum_{sy,sx in {0,1}} in[y+sy, x+sx, :] . Wt[:, co, ky, kx],  ky = py+1-2sy, kx = px+1-2sx (if in 0..2)
//   Wbf[s][(p,co)][ci]  (forward, N = 4Co, Kt = Ci)      Wbd[s][ci][(p,co)]  (input gradient, N = Ci, Kt = 4Co)
__global__ void k_prep_ct_w(const float *__restrict__ params, VvIntG slot, long long slot_stride, long long w_off, long long b_off,
                            int Ci, int Co, float *__restrict__ Wbf, long long wf_gs, float *__restrict__ Wbd, long long wd_gs,
                            float *__restrict__ vec, long long vec_gs) {
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
        Wbf[g * wf_gs + i] = v;                                                     // [s][ph*Co+co][ci]
        Wbd[g * wd_gs + ((long long)s * Ci + ci) * (4 * Co) + ph * Co + co] = v;    // [s][ci][ph*Co+co]
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

// thread = (output pixel or pooled pixel, 4 channels).  blockDim = (C/4 <= 128 .. , rows)
__global__ void k_bn_apply(const VvBnApply p) {
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
    const float *Z = p.Z + g * p.z_gs;
    float *Y = p.Y + g * p.y_gs;
    const int cq = p.C >> 2;
    if (!p.pool) {
        for (int c4 = threadIdx.x; c4 < cq; c4 += blockDim.x) {
            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c4 * 4);
            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c4 * 4);
            for (int m = blockIdx.x * blockDim.y + threadIdx.y; m < p.M; m += gridDim.x * blockDim.y) {
                float4 z = *reinterpret_cast<const float4 *>(Z + (long long)m * p.C + c4 * 4);
                float4 y;
                y.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f); y.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
                y.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f); y.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
                *reinterpret_cast<float4 *>(Y + (long long)m * p.ldy + p.y_coff + c4 * 4) = y;
            }
        }
    } else {
        float *Pl = p.P + g * p.p_gs;
        const int Hp = p.H >> 1, Wp = p.W >> 1;
        const int Mp = p.M >> 2;
        for (int c4 = threadIdx.x; c4 < cq; c4 += blockDim.x) {
            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c4 * 4);
            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c4 * 4);
            for (int mp = blockIdx.x * blockDim.y + threadIdx.y; mp < Mp; mp += gridDim.x * blockDim.y) {
                int b, yp, xp;
                pix3(mp, Hp, Wp, b, yp, xp);
                float4 mx = make_float4(0.f, 0.f, 0.f, 0.f);   // post-ReLU values are >= 0
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    long long m = ((long long)(b * p.H + 2 * yp + (w >> 1))) * p.W + 2 * xp + (w & 1);
                    float4 z = *reinterpret_cast<const float4 *>(Z + m * p.C + c4 * 4);
                    float4 y;
                    y.x = fmaxf(fmaf(z.x, sc.x, sh.x), 0.f); y.y = fmaxf(fmaf(z.y, sc.y, sh.y), 0.f);
                    y.z = fmaxf(fmaf(z.z, sc.z, sh.z), 0.f); y.w = fmaxf(fmaf(z.w, sc.w, sh.w), 0.f);
                    *reinterpret_cast<float4 *>(Y + m * p.ldy + p.y_coff + c4 * 4) = y;
                    mx.x = fmaxf(mx.x, y.x); mx.y = fmaxf(mx.y, y.y); mx.z = fmaxf(mx.z, y.z); mx.w = fmaxf(mx.w, y.w);
                }
                *reinterpret_cast<float4 *>(Pl + (long long)mp * p.C + c4 * 4) = mx;
            }
        }
    }
}

// ---- BatchNorm backward through ReLU.  dzhat = dy * [z*scale+shift > 0];  xhat = (z-mean)*invstd
//      pass 1: sums[g][0][c] = sum dzhat, sums[g][1][c] = sum dzhat*xhat  (double atomics)
__global__ void k_bn_bwd_reduce(const VvBnBwd p) {
    extern __shared__ float sm[];   // [2][blockDim.y][C] partials
    const int g = blockIdx.y;
    const float *Z = p.Z + g * p.z_gs;
    const float *dY = p.dY + g * p.dy_gs;
    const float *sv = p.save + g * p.save_gs;
    const int cq = p.C >> 2;
    float *ps = sm, *pq = sm + blockDim.y * p.C;
    for (int c4 = threadIdx.x; c4 < cq; c4 += blockDim.x) {
        const float4 sc = *reinterpret_cast<const float4 *>(sv + c4 * 4);
        const float4 sh = *reinterpret_cast<const float4 *>(sv + p.C + c4 * 4);
        const float4 mu = *reinterpret_cast<const float4 *>(sv + 2 * p.C + c4 * 4);
        const float4 is = *reinterpret_cast<const float4 *>(sv + 3 * p.C + c4 * 4);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
        for (int m = blockIdx.x * blockDim.y + threadIdx.y; m < p.M; m += gridDim.x * blockDim.y) {
            float4 z = *reinterpret_cast<const float4 *>(Z + (long long)m * p.C + c4 * 4);
            float4 d = *reinterpret_cast<const float4 *>(dY + (long long)m * p.ldy + p.dy_coff + c4 * 4);
            float dx = fmaf(z.x, sc.x, sh.x) > 0.f ? d.x : 0.f;
            float dy = fmaf(z.y, sc.y, sh.y) > 0.f ? d.y : 0.f;
            float dz = fmaf(z.z, sc.z, sh.z) > 0.f ? d.z : 0.f;
            float dw = fmaf(z.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
            s.x += dx; s.y += dy; s.z += dz; s.w += dw;
            q.x += dx * (z.x - mu.x) * is.x; q.y += dy * (z.y - mu.y) * is.y;
            q.z += dz * (z.z - mu.z) * is.z; q.w += dw * (z.w - mu.w) * is.w;
        }
        *reinterpret_cast<float4 *>(ps + threadIdx.y * p.C + c4 * 4) = s;
        *reinterpret_cast<float4 *>(pq + threadIdx.y * p.C + c4 * 4) = q;
    }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y;
    double *sums = p.sums + g * p.sums_gs;
    for (int c = tid; c < p.C; c += nthr) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < blockDim.y; r++) { a += ps[r * p.C + c]; b += pq[r * p.C + c]; }
        atomicAdd(&sums[c], (double)a);
        atomicAdd(&sums[p.C + c], (double)b);
    }
}

//      pass 2: dz = scale * (dzhat - mean(dzhat) - xhat * mean(dzhat*xhat));  d gamma = sum dzhat*xhat, d beta = sum dzhat
__global__ void k_bn_bwd_apply(const VvBnBwd p) {
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
            G[p.gamma_off + c] = (float)b;
            G[p.beta_off + c] = (float)a;
        }
    }
    __syncthreads();
    const float *Z = p.Z + g * p.z_gs;
    const float *dY = p.dY + g * p.dy_gs;
    float *dZ = p.dZ + g * p.dz_gs;
    const float *sv = p.save + g * p.save_gs;
    const int cq = p.C >> 2;
    for (int c4 = threadIdx.x; c4 < cq; c4 += blockDim.x) {
        const float4 sc = *reinterpret_cast<const float4 *>(sv + c4 * 4);
        const float4 sh = *reinterpret_cast<const float4 *>(sv + p.C + c4 * 4);
        const float4 mu = *reinterpret_cast<const float4 *>(sv + 2 * p.C + c4 * 4);
        const float4 is = *reinterpret_cast<const float4 *>(sv + 3 * p.C + c4 * 4);
        const float4 a1 = *reinterpret_cast<const float4 *>(k1 + c4 * 4);
        const float4 a2 = *reinterpret_cast<const float4 *>(k2 + c4 * 4);
        for (int m = blockIdx.x * blockDim.y + threadIdx.y; m < p.M; m += gridDim.x * blockDim.y) {
            float4 z = *reinterpret_cast<const float4 *>(Z + (long long)m * p.C + c4 * 4);
            float4 d = *reinterpret_cast<const float4 *>(dY + (long long)m * p.ldy + p.dy_coff + c4 * 4);
            float4 o;
            o.x = sc.x * ((fmaf(z.x, sc.x, sh.x) > 0.f ? d.x : 0.f) - a1.x - (z.x - mu.x) * is.x * a2.x);
            o.y = sc.y * ((fmaf(z.y, sc.y, sh.y) > 0.f ? d.y : 0.f) - a1.y - (z.y - mu.y) * is.y * a2.y);
            o.z = sc.z * ((fmaf(z.z, sc.z, sh.z) > 0.f ? d.z : 0.f) - a1.z - (z.z - mu.z) * is.z * a2.z);
            o.w = sc.w * ((fmaf(z.w, sc.w, sh.w) > 0.f ? d.w : 0.f) - a1.w - (z.w - mu.w) * is.w * a2.w);
            *reinterpret_cast<float4 *>(dZ + (long long)m * p.C + c4 * 4) = o;
        }
    }
}

// ---- MaxPool2d(2) backward: the gradient of each pooled element is added to the FIRST maximum of its 2x2 window in
//      row-major order (ATen max_pool2d keeps the first index on ties, which are frequent after ReLU).
__global__ void k_maxpool_bwd(const float *__restrict__ Y, long long y_gs, int ldy, int y_coff, const float *__restrict__ dP,
                              long long dp_gs, float *__restrict__ dY, long long dy_gs, int lddy, int dy_coff, int B, int H, int W,
                              int C) {
    const int g = blockIdx.y;
    const int Hp = H >> 1, Wp = W >> 1, cq = C >> 2;
    const long long total = (long long)B * Hp * Wp * cq;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % cq);
        int mp = (int)(i / cq);
        int b, yp, xp;
        pix3(mp, Hp, Wp, b, yp, xp);
        float4 dp = *reinterpret_cast<const float4 *>(dP + g * dp_gs + (long long)mp * C + c4 * 4);
        float4 v[4];
        long long mi[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            mi[w] = ((long long)(b * H + 2 * yp + (w >> 1))) * W + 2 * xp + (w & 1);
            v[w] = *reinterpret_cast<const float4 *>(Y + g * y_gs + mi[w] * ldy + y_coff + c4 * 4);
        }
        int ax = 0, ay = 0, az = 0, aw = 0;
        float bx = v[0].x, by = v[0].y, bz = v[0].z, bw = v[0].w;
#pragma unroll
        for (int w = 1; w < 4; w++) {
            if (v[w].x > bx) { bx = v[w].x; ax = w; }
            if (v[w].y > by) { by = v[w].y; ay = w; }
            if (v[w].z > bz) { bz = v[w].z; az = w; }
            if (v[w].w > bw) { bw = v[w].w; aw = w; }
        }
#pragma unroll
        for (int w = 0; w < 4; w++) {
            float *d = dY + g * dy_gs + mi[w] * lddy + dy_coff + c4 * 4;
            float4 cur = *reinterpret_cast<float4 *>(d);
            cur.x += (ax == w) ? dp.x : 0.f; cur.y += (ay == w) ? dp.y : 0.f;
            cur.z += (az == w) ? dp.z : 0.f; cur.w += (aw == w) ? dp.w : 0.f;
            *reinterpret_cast<float4 *>(d) = cur;
        }
    }
}

// ---- column sum of a (strided) gradient tensor -> ConvTranspose2d bias gradient
__global__ void k_colsum(const float *__restrict__ D, long long d_gs, int ld, int coff, int M, int C, float *__restrict__ grads,
                         VvIntG slot, long long slot_stride, long long off) {
    extern __shared__ float sm[];   // [blockDim.y][C]
    const int g = blockIdx.y;
    const int cq = C >> 2;
    for (int c4 = threadIdx.x; c4 < cq; c4 += blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = blockIdx.x * blockDim.y + threadIdx.y; m < M; m += gridDim.x * blockDim.y) {
            float4 d = *reinterpret_cast<const float4 *>(D + g * d_gs + (long long)m * ld + coff + c4 * 4);
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
        }
        *reinterpret_cast<float4 *>(sm + threadIdx.y * C + c4 * 4) = s;
    }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x * blockDim.y) {
        float a = 0.f;
        for (int r = 0; r < blockDim.y; r++) a += sm[r * C + c];
        atomicAdd(grads + slot.v[g] * slot_stride + off + c, a);
    }
}

// ---- 1x1 output conv (model/unet.py:63-70) with fused squared error and MSE gradient (train.py:385-392, 414-427).
//      One CTA = 256 threads = one cube (S*S pixels, S*S/256 pixels per thread) of one UNet: the per-cube SSE is a
//      deterministic in-CTA reduction.  U [G][B*S*S][F];  out NCHW;  dout [G][B*S*S][4].
__global__ void __launch_bounds__(256) k_outconv_fwd(const VvOutFwd p) {
    extern __shared__ float sm[];   // tile [256][F+1], w [4][F], b[4]
    const int g = blockIdx.y, b = blockIdx.x;
    const int F = p.F, SS = p.S * p.S;
    float *tile = sm;
    float *w = sm + 256 * (F + 1);
    float *bs = w + 4 * F;
    __shared__ float red[8];
    const int oc = p.out_channels.v[g];
    const float *P = p.params + p.slot.v[g] * p.slot_param_stride;
    for (int i = threadIdx.x; i < 4 * F; i += 256) w[i] = (i < oc * F) ? P[p.w_off + i] : 0.f;
    if (threadIdx.x < 4) bs[threadIdx.x] = threadIdx.x < oc ? P[p.b_off + threadIdx.x] : 0.f;
    const float *U = p.U + g * p.u_gs + (long long)b * SS * F;
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
    for (int base = 0; base < SS; base += 256) {
        __syncthreads();
        // coalesced load of 256 pixels x F channels
        for (int i = threadIdx.x; i < 256 * F / 4; i += 256) {
            int r = (i * 4) / F, c = (i * 4) % F;
            float4 v = *reinterpret_cast<const float4 *>(U + (long long)(base + r) * F + c);
            float *d = tile + r * (F + 1) + c;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        const int pix = base + threadIdx.x;
        float o[4] = {bs[0], bs[1], bs[2], bs[3]};
        const float *row = tile + threadIdx.x * (F + 1);
        for (int c = 0; c < F; c++) {
            float u = row[c];
            o[0] = fmaf(u, w[c], o[0]); o[1] = fmaf(u, w[F + c], o[1]); o[2] = fmaf(u, w[2 * F + c], o[2]);
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
// warp = 32 channels of one pixel at a time (lane = channel), grid-stride over pixels; F is processed in slabs of 32.
__global__ void __launch_bounds__(256) k_outconv_bwd(const VvOutBwd p) {
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = p.F;
    const int oc = p.out_channels.v[g];
    const float *P = p.params + p.slot.v[g] * p.slot_param_stride;
    float *G = p.grads + p.slot.v[g] * p.slot_param_stride;
    const float *U = p.U + g * p.u_gs;
    float *dU = p.dU + g * p.du_gs;
    const int SS = p.S * p.S;
    // external NCHW gradient or the internal staged [m][4]
    const bool flow = p.target_is_flow.v[g] != 0;
    const float *ext = flow ? p.grad_of_out : p.grad_raw_out;
    const int ext_ctot = flow ? p.of_out_channels : p.raw_out_channels;
    const int ext_c0 = p.out_slot.v[g] * (flow ? 2 : 3);
    __shared__ float red[8][3][32];
    __shared__ float redb[8][4];
    for (int f0 = 0; f0 < F; f0 += 32) {
        float w0 = oc > 0 ? P[p.w_off + 0 * F + f0 + lane] : 0.f;
        float w1 = oc > 1 ? P[p.w_off + 1 * F + f0 + lane] : 0.f;
        float w2 = oc > 2 ? P[p.w_off + 2 * F + f0 + lane] : 0.f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, db = 0.f;
        for (int m = blockIdx.x * 8 + warp; m < p.M; m += gridDim.x * 8) {
            float d0, d1, d2;
            if (ext) {
                int b = m / SS, pix = m - b * SS;
                const float *e = ext + ((long long)b * ext_ctot + ext_c0) * SS + pix;
                d0 = e[0];
                d1 = oc > 1 ? e[SS] : 0.f;
                d2 = oc > 2 ? e[2 * SS] : 0.f;
            } else {
                float4 d = *reinterpret_cast<const float4 *>(p.dout + ((long long)g * p.M + m) * 4);
                d0 = d.x; d1 = d.y; d2 = d.z;
            }
            float u = U[(long long)m * F + f0 + lane];
            a0 = fmaf(d0, u, a0); a1 = fmaf(d1, u, a1); a2 = fmaf(d2, u, a2);
            dU[(long long)m * F + f0 + lane] = d0 * w0 + d1 * w1 + d2 * w2;
            if (f0 == 0) db += (lane == 0) ? d0 : (lane == 1) ? d1 : (lane == 2) ? d2 : 0.f;
        }
        red[warp][0][lane] = a0; red[warp][1][lane] = a1; red[warp][2][lane] = a2;
        if (lane < 4) redb[warp][lane] = db;
        __syncthreads();
        if (threadIdx.x < 96) {
            int j = threadIdx.x >> 5;
            float s = 0.f;
            for (int wv = 0; wv < 8; wv++) s += red[wv][j][lane];
            if (j < oc) atomicAdd(G + p.w_off + j * F + f0 + lane, s);
        }
        if (f0 == 0 && threadIdx.x >= 96 && threadIdx.x < 96 + 3) {
            int j = threadIdx.x - 96;
            float s = 0.f;
            for (int wv = 0; wv < 8; wv++) s += redb[wv][j];
            if (j < oc) atomicAdd(G + p.b_off + j, s);
        }
        __syncthreads();
    }
}

// ---- gradient re-layout back to PyTorch's parameter layouts
__global__ void k_scatter_conv_wgrad(const float *__restrict__ dWf, long long gs, int N, int C, int Cp, float *__restrict__ grads,
                                     VvIntG slot, long long slot_stride, long long w_off) {
    const int g = blockIdx.y;
    const int total = N * C * 9;
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // i indexes the PyTorch layout [n][c][t] (coalesced writes)
    if (i >= total) return;
    int t = i % 9, c = (i / 9) % C, n = i / (9 * C);
    grads[slot.v[g] * slot_stride + w_off + i] = dWf[g * gs + ((long long)t * N + n) * Cp + c];
}

__global__ void k_scatter_ct_wgrad(const float *__restrict__ dWb, long long gs, int Ci, int Co, float *__restrict__ grads,
                                   VvIntG slot, long long slot_stride, long long w_off) {
    const int g = blockIdx.y;
    const int total = Ci * Co * 9;
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // PyTorch layout [ci][co][ky][kx]
    if (i >= total) return;
    int kx = i % 3, ky = (i / 3) % 3, co = (i / 9) % Co, ci = i / (9 * Co);
    // ky = py + 1 - 2 sy  ->  ky=0:(py=1,sy=1)  ky=1:(py=0,sy=0)  ky=2:(py=1,sy=0)
    int py = (ky == 1) ? 0 : 1, sy = (ky == 0) ? 1 : 0;
    int px = (kx == 1) ? 0 : 1, sx = (kx == 0) ? 1 : 0;
    int s = sy * 2 + sx, ph = py * 2 + px;
    grads[slot.v[g] * slot_stride + w_off + i] = dWb[g * gs + ((long long)s * 4 * Co + ph * Co + co) * Ci + ci];
}

// ---- losses from the per-cube SSE buffer: mean over B * ch * S * S (train.py:385-392)
__global__ void k_losses(const float *__restrict__ sse, int G, int B, VvIntG is_flow, float inv_raw, float inv_of, float *__restrict__ out) {
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
__global__ void k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, long long n,
                       float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
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

static inline dim3 row_block(int C, int &rows) {
    int tx = C / 4;
    if (tx > 64) tx = 64;
    rows = 256 / tx;
    return dim3(tx, rows);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int vv_prep_input(const float *x, float *X0, int G, int B, int T, int S, int cinp, int padding, const VvIntG &erase, cudaStream_t st) {
    int M = B * S * S;
    k_prep_input<<<dim3(vv_cdiv(M, 256), G), 256, 0, st>>>(x, X0, B, T, S, cinp, padding, erase);
    VV_CKL();
    return 0;
}

int vv_prep_conv_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, long long g_off,
                   long long beta_off, int N, int C, int Cp, float *Wf, long long wf_gs, float *Wd, long long wd_gs, float *vec,
                   long long vec_gs, int G, cudaStream_t st) {
    k_prep_conv_w<<<dim3(vv_cdiv(9LL * N * Cp, 256), G), 256, 0, st>>>(params, slot, slot_stride, w_off, b_off, g_off, beta_off, N, C,
                                                                     Cp, Wf, wf_gs, Wd, wd_gs, vec, vec_gs);
    VV_CKL();
    return 0;
}

int vv_prep_ct_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, int Ci, int Co,
                 float *Wbf, long long wf_gs, float *Wbd, long long wd_gs, float *vec, long long vec_gs, int G, cudaStream_t st) {
    k_prep_ct_w<<<dim3(vv_cdiv(16LL * Co * Ci, 256), G), 256, 0, st>>>(params, slot, slot_stride, w_off, b_off, Ci, Co, Wbf, wf_gs, Wbd,
                                                                     wd_gs, vec, vec_gs);
    VV_CKL();
    return 0;
}

int vv_bn_apply(const VvBnApply &p, int G, cudaStream_t st) {
    int rows;
    dim3 blk = row_block(p.C, rows);
    int work = p.pool ? p.M / 4 : p.M;
    int gx = vv_cdiv(work, rows * 4);
    if (gx > 148 * 8) gx = 148 * 8;
    if (gx < 1) gx = 1;
    k_bn_apply<<<dim3(gx, G), blk, 2 * p.C * sizeof(float), st>>>(p);
    VV_CKL();
    return 0;
}

int vv_bn_bwd(const VvBnBwd &p, int G, cudaStream_t st) {
    int rows;
    dim3 blk = row_block(p.C, rows);
    int gx = vv_cdiv(p.M, rows * 8);
    if (gx > 148 * 4) gx = 148 * 4;
    if (gx < 1) gx = 1;
    k_bn_bwd_reduce<<<dim3(gx, G), blk, 2 * rows * p.C * sizeof(float), st>>>(p);
    VV_CKL();
    int gx2 = vv_cdiv(p.M, rows * 4);
    if (gx2 > 148 * 8) gx2 = 148 * 8;
    if (gx2 < 1) gx2 = 1;
    k_bn_bwd_apply<<<dim3(gx2, G), blk, 2 * p.C * sizeof(float), st>>>(p);
    VV_CKL();
    return 0;
}

int vv_maxpool_bwd(const float *Y, long long y_gs, int ldy, int y_coff, const float *dP, long long dp_gs, float *dY, long long dy_gs,
                   int lddy, int dy_coff, int G, int B, int H, int W, int C, cudaStream_t st) {
    long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
    int gx = vv_cdiv(total, 256);
    if (gx > 148 * 16) gx = 148 * 16;
    k_maxpool_bwd<<<dim3(gx, G), 256, 0, st>>>(Y, y_gs, ldy, y_coff, dP, dp_gs, dY, dy_gs, lddy, dy_coff, B, H, W, C);
    VV_CKL();
    return 0;
}

int vv_colsum(const float *D, long long d_gs, int ld, int coff, int M, int C, float *grads, const VvIntG &slot, long long slot_stride,
              long long off, int G, cudaStream_t st) {
    int rows;
    dim3 blk = row_block(C, rows);
    int gx = vv_cdiv(M, rows * 16);
    if (gx > 148 * 2) gx = 148 * 2;
    if (gx < 1) gx = 1;
    k_colsum<<<dim3(gx, G), blk, rows * C * sizeof(float), st>>>(D, d_gs, ld, coff, M, C, grads, slot, slot_stride, off);
    VV_CKL();
    return 0;
}

int vv_outconv_fwd(const VvOutFwd &p, int G, cudaStream_t st) {
    VV_REQUIRE((p.S * p.S) % 256 == 0, "outconv: S*S must be a multiple of 256 (S=%d)", p.S);
    size_t smem = (256 * (p.F + 1) + 4 * p.F + 4) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        VV_CK(cudaFuncSetAttribute(k_outconv_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    k_outconv_fwd<<<dim3(p.B, G), 256, smem, st>>>(p);
    VV_CKL();
    return 0;
}

int vv_outconv_bwd(const VvOutBwd &p, int G, cudaStream_t st) {
    VV_REQUIRE(p.F % 32 == 0, "outconv backward needs features_root %% 32 == 0");
    int gx = vv_cdiv(p.M, 8 * 32);
    if (gx > 148 * 4) gx = 148 * 4;
    k_outconv_bwd<<<dim3(gx, G), 256, 0, st>>>(p);
    VV_CKL();
    return 0;
}

int vv_scatter_conv_wgrad(const float *dWf, long long gs, int N, int C, int Cp, float *grads, const VvIntG &slot, long long slot_stride,
                          long long w_off, int G, cudaStream_t st) {
    k_scatter_conv_wgrad<<<dim3(vv_cdiv(9LL * N * C, 256), G), 256, 0, st>>>(dWf, gs, N, C, Cp, grads, slot, slot_stride, w_off);
    VV_CKL();
    return 0;
}

int vv_scatter_ct_wgrad(const float *dWb, long long gs, int Ci, int Co, float *grads, const VvIntG &slot, long long slot_stride,
                        long long w_off, int G, cudaStream_t st) {
    k_scatter_ct_wgrad<<<dim3(vv_cdiv(9LL * Ci * Co, 256), G), 256, 0, st>>>(dWb, gs, Ci, Co, grads, slot, slot_stride, w_off);
    VV_CKL();
    return 0;
}

int vv_losses(const float *sse, int G, int B, const VvIntG &is_flow, float inv_raw, float inv_of, float *out, cudaStream_t st) {
    k_losses<<<1, 256, 0, st>>>(sse, G, B, is_flow, inv_raw, inv_of, out);
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
    k_adam<<<vv_cdiv(nthr, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                               (float)bc1, (float)sqrt(bc2), grad_scale);
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
    k_cubes_to_tensors<<<gx, 256, 0, (cudaStream_t)stream>>>(raw, flow, x, x_of, n, t_raw, t_of, patch);
    VV_CKL();
    return 0;
}
