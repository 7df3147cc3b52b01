// fp32 SIMT implicit-GEMM tiles: exact-fp32 path of the conv / transposed-conv / dgrad / wgrad
// contractions of the completion UNets (reference: cuDNN calls behind model/unet.py:9-16,54,66).
// Used when vecvad_net_config.use_tensor_cores == 0, for shapes the tcgen05 tiles do not cover
// (ragged batches), and as the on-device fp32 cross-check of the tcgen05 tiles in tests.
#include "common.h"

namespace {

constexpr int BM = 128;   // pixels per CTA tile
constexpr int BK = 16;    // K chunk
constexpr int AS = BM + 4;

__device__ __forceinline__ void pix_decompose(int m, int H, int W, int &b, int &y, int &x) {
    x = m % W;
    int t = m / W;
    y = t % H;
    b = t / H;
}

// address of A[(b,yy,xx)][k] for plain / space-to-depth sources (k multiple of 4; 4 consecutive k never straddle a phase)
__device__ __forceinline__ const float *a_addr(const float *A, int lda, int coff, int s2d, int Kt, int H, int W, int b, int yy,
                                               int xx, int k) {
    if (!s2d) return A + ((long long)(b * H + yy) * W + xx) * lda + coff + k;
    int cq = Kt >> 2;
    int ph = k / cq, c = k - ph * cq;
    int py = ph >> 1, px = ph & 1;
    return A + ((long long)(b * 2 * H + 2 * yy + py) * (2 * W) + 2 * xx + px) * lda + coff + c;
}

// ------------------------------------------------------------------------------------------------
// out[m,n] = bias + sum_t sum_k A[shift(m,t),k] * Wt[t][n][k]        TN = columns per thread (BN = 16*TN)
// ------------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(256) k_igemm_simt(const VvIGemm p) {
    vv_pdl_wait();
    constexpr int BN = 16 * TN;
    constexpr int BS = BN + 4;
    __shared__ __align__(16) float As[2][BK][AS];
    __shared__ __align__(16) float Bs[2][BK][BS];
    __shared__ float s_sum[BN], s_sq[BN];

    const int g = blockIdx.z;
    const int M = p.B * p.H * p.W;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int t = threadIdx.x;
    const float *A = p.A + g * p.a_gs;
    const float *Wt = p.Wt + g * p.w_gs;

    // A-load role: one pixel row, 8 consecutive k
    const int arow = t & (BM - 1);
    const int akh = (t >> 7) * 8;
    int ab, ay, ax;
    const int am = m0 + arow;
    pix_decompose(am < M ? am : 0, p.H, p.W, ab, ay, ax);
    // B-load role: TN==4: 64 n x 16 k = 256 float4 ; TN==2: 32 n x 16 k = 128 float4 (threads >= 128 idle)
    const int bn = t >> 2;
    const int bkq = (t & 3) * 4;
    const bool b_active = (bn < BN) && (n0 + bn < p.N);

    const int kc = p.Kt / BK;
    const int nchunks = p.taps.n * kc;

    float4 ra0, ra1, rb;
    auto load_regs = [&](int chunk) {
        int tap = chunk / kc;
        int c0 = (chunk - tap * kc) * BK;
        int yy = ay + p.taps.dy[tap], xx = ax + p.taps.dx[tap];
        bool ok = (am < M) && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        if (ok) {
            const float *src = a_addr(A, p.lda, p.a_coff, p.a_s2d, p.Kt, p.H, p.W, ab, yy, xx, c0 + akh);
            ra0 = *reinterpret_cast<const float4 *>(src);
            if (!p.a_s2d) {
                ra1 = *reinterpret_cast<const float4 *>(src + 4);
            } else {
                ra1 = *reinterpret_cast<const float4 *>(
                    a_addr(A, p.lda, p.a_coff, p.a_s2d, p.Kt, p.H, p.W, ab, yy, xx, c0 + akh + 4));
            }
        } else {
            ra0 = make_float4(0.f, 0.f, 0.f, 0.f);
            ra1 = ra0;
        }
        if (b_active)
            rb = *reinterpret_cast<const float4 *>(Wt + ((long long)tap * p.N + n0 + bn) * p.Kt + c0 + bkq);
        else
            rb = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_smem = [&](int buf) {
        As[buf][akh + 0][arow] = ra0.x; As[buf][akh + 1][arow] = ra0.y; As[buf][akh + 2][arow] = ra0.z; As[buf][akh + 3][arow] = ra0.w;
        As[buf][akh + 4][arow] = ra1.x; As[buf][akh + 5][arow] = ra1.y; As[buf][akh + 6][arow] = ra1.z; As[buf][akh + 7][arow] = ra1.w;
        if (bn < BN) {
            Bs[buf][bkq + 0][bn] = rb.x; Bs[buf][bkq + 1][bn] = rb.y; Bs[buf][bkq + 2][bn] = rb.z; Bs[buf][bkq + 3][bn] = rb.w;
        }
    };

    const int tx = t & 15, ty = t >> 4;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    load_regs(0);
    store_smem(0);
    __syncthreads();
    for (int ch = 0; ch < nchunks; ch++) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) load_regs(ch + 1);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
            if (TN == 4) {
                float4 b4 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
                bv[0] = b4.x; bv[1] = b4.y; bv[TN - 2] = b4.z; bv[TN - 1] = b4.w;
            } else {
                float2 b2 = *reinterpret_cast<const float2 *>(&Bs[buf][k][tx * 2]);
                bv[0] = b2.x; bv[1] = b2.y;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (ch + 1 < nchunks) store_smem(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, store, per-channel statistics ----
    const int ncol = n0 + tx * TN;
    const bool col_ok = ncol < p.N;
    const int Co = p.o_d2s ? (p.N >> 2) : p.N;
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; j++) bv[j] = 0.f;
    if (p.bias && col_ok) {
        const float *bias = p.bias + g * p.bias_gs;
#pragma unroll
        for (int j = 0; j < TN; j++) bv[j] = bias[(ncol + j) % Co];
    }
    float cs[TN], cq[TN];
#pragma unroll
    for (int j = 0; j < TN; j++) { cs[j] = 0.f; cq[j] = 0.f; }
    float *O = p.O + g * p.o_gs;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int m = m0 + ty * 8 + i;
        if (m < M && col_ok) {
            float v[TN];
#pragma unroll
            for (int j = 0; j < TN; j++) {
                v[j] = acc[i][j] + bv[j];
                cs[j] += v[j];
                cq[j] += v[j] * v[j];
            }
            float *dst;
            if (!p.o_d2s) {
                dst = O + (long long)m * p.ldo + p.o_coff + ncol;
            } else {
                int b, y, x;
                pix_decompose(m, p.H, p.W, b, y, x);
                int ph = ncol / Co, co = ncol - ph * Co;
                dst = O + ((long long)(b * 2 * p.H + 2 * y + (ph >> 1)) * (2 * p.W) + 2 * x + (ph & 1)) * p.ldo + p.o_coff + co;
            }
            if (TN == 4)
                *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[TN - 2], v[TN - 1]);
            else
                *reinterpret_cast<float2 *>(dst) = make_float2(v[0], v[1]);
        }
    }
    if (p.stats) {
        if (t < BN) { s_sum[t] = 0.f; s_sq[t] = 0.f; }
        __syncthreads();
        if (col_ok) {
#pragma unroll
            for (int j = 0; j < TN; j++) {
                atomicAdd(&s_sum[tx * TN + j], cs[j]);
                atomicAdd(&s_sq[tx * TN + j], cq[j]);
            }
        }
        __syncthreads();
        if (t < BN && n0 + t < p.N) {
            double *st = p.stats + g * p.stats_gs;
            atomicAdd(&st[n0 + t], (double)s_sum[t]);
            atomicAdd(&st[p.N + n0 + t], (double)s_sq[t]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dW[t][n][k] += sum_m Gd[m,n] * A[shift(m,t),k]
// CTA tile: 64 n  x  4 "sub-blocks" of 16 k (each sub-block has its own tap), reduced over a slice of pixels.
// ------------------------------------------------------------------------------------------------
constexpr int WK = 16;  // pixels per step
__global__ void __launch_bounds__(256) k_wgrad_simt(const VvWGrad p, int nqt, int rows_per_split) {
    vv_pdl_wait();
    __shared__ __align__(16) float Gs[2][WK][64 + 4];
    __shared__ __align__(16) float As[2][WK][64 + 4];
    const int g = blockIdx.z;
    const int M = p.B * p.H * p.W;
    const int ntile = blockIdx.x / nqt, qtile = blockIdx.x - ntile * nqt;
    const int n0 = ntile * 64;
    const int kc = p.Kt / 16;
    const int nsub = p.taps.n * kc;
    const int t = threadIdx.x;
    const float *A = p.A + g * p.a_gs;
    const float *Gd = p.Gd + g * p.g_gs;

    const int mbeg = blockIdx.y * rows_per_split;
    const int mend = min(M, mbeg + rows_per_split);
    if (mbeg >= mend) return;

    // load roles
    const int lk = t >> 4;            // pixel within step
    const int lq = t & 15;            // float4 index within the 64-wide row
    // G operand
    const int gn = n0 + lq * 4;
    const bool g_ok = gn < p.N;
    // A operand
    const int sub = qtile * 4 + (lq >> 2);
    const bool a_ok = sub < nsub;
    const int a_tap = a_ok ? sub / kc : 0;
    const int a_c = a_ok ? (sub - a_tap * kc) * 16 + (lq & 3) * 4 : 0;
    const int a_dy = p.taps.dy[a_tap], a_dx = p.taps.dx[a_tap];

    float4 rg, ra;
    auto load_regs = [&](int ms) {
        int m = ms + lk;
        rg = make_float4(0.f, 0.f, 0.f, 0.f);
        ra = rg;
        if (m < mend) {
            int b, y, x;
            pix_decompose(m, p.H, p.W, b, y, x);
            if (g_ok) rg = *reinterpret_cast<const float4 *>(a_addr(Gd, p.ldg, p.g_coff, p.g_s2d, p.N, p.H, p.W, b, y, x, gn));
            int yy = y + a_dy, xx = x + a_dx;
            if (a_ok && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
                ra = *reinterpret_cast<const float4 *>(A + ((long long)(b * p.H + yy) * p.W + xx) * p.lda + p.a_coff + a_c);
        }
    };
    auto store_smem = [&](int buf) {
        *reinterpret_cast<float4 *>(&Gs[buf][lk][lq * 4]) = rg;
        *reinterpret_cast<float4 *>(&As[buf][lk][lq * 4]) = ra;
    };

    const int tx = t & 15, ty = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    load_regs(mbeg);
    store_smem(0);
    __syncthreads();
    int it = 0;
    for (int ms = mbeg; ms < mend; ms += WK, it++) {
        const int buf = it & 1;
        const bool more = ms + WK < mend;
        if (more) load_regs(ms + WK);
#pragma unroll
        for (int k = 0; k < WK; k++) {
            float4 gv = *reinterpret_cast<const float4 *>(&Gs[buf][k][ty * 4]);
            float4 av = *reinterpret_cast<const float4 *>(&As[buf][k][tx * 4]);
            float gg[4] = {gv.x, gv.y, gv.z, gv.w};
            float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(gg[i], aa[j], acc[i][j]);
        }
        if (more) store_smem(buf ^ 1);
        __syncthreads();
    }
    // epilogue
    const int osub = qtile * 4 + (tx >> 2);
    if (osub < nsub) {
        const int otap = osub / kc;
        const int oc = (osub - otap * kc) * 16 + (tx & 3) * 4;
        float *dW = p.dW + g * p.dw_gs;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int n = n0 + ty * 4 + i;
            if (n < p.N) {
                float *dst = dW + ((long long)otap * p.N + n) * p.Kt + oc;
#pragma unroll
                for (int j = 0; j < 4; j++) atomicAdd(dst + j, acc[i][j]);
            }
        }
    }
}

}  // namespace

int vv_launch_igemm_simt(const VvIGemm &p, cudaStream_t st) {
    VV_REQUIRE(p.Kt % 16 == 0 && p.N % 16 == 0, "igemm: Kt=%d N=%d must be multiples of 16", p.Kt, p.N);
    VV_REQUIRE(!p.a_s2d || (p.Kt % 64 == 0), "igemm: s2d source needs Kt %% 64 == 0 (Kt=%d)", p.Kt);
    VV_REQUIRE(!p.o_d2s || (p.N % 16 == 0 && (p.N / 4) % 4 == 0), "igemm: d2s output needs N/4 %% 4 == 0");
    int M = p.B * p.H * p.W;
    if (p.N % 64 == 0 || p.N > 64) {
        dim3 grid(vv_cdiv(M, BM), vv_cdiv(p.N, 64), p.G);
        vv_launch(k_igemm_simt<4>, dim3(grid), dim3(256), 0, st, p);
    } else {
        dim3 grid(vv_cdiv(M, BM), vv_cdiv(p.N, 32), p.G);
        vv_launch(k_igemm_simt<2>, dim3(grid), dim3(256), 0, st, p);
    }
    VV_CKL();
    return 0;
}

int vv_launch_wgrad_simt(const VvWGrad &p, cudaStream_t st) {
    VV_REQUIRE(p.Kt % 16 == 0 && p.N % 16 == 0, "wgrad: Kt=%d N=%d must be multiples of 16", p.Kt, p.N);
    VV_REQUIRE(!p.g_s2d || ((p.N / 4) % 4 == 0), "wgrad: s2d gradient needs N/4 %% 4 == 0");
    int M = p.B * p.H * p.W;
    int nsub = p.taps.n * (p.Kt / 16);
    int nqt = vv_cdiv(nsub, 4);
    int nnt = vv_cdiv(p.N, 64);
    long long tiles = (long long)nqt * nnt * p.G;
    int msplit = (int)((148LL * 8 + tiles - 1) / tiles);
    int max_split = vv_cdiv(M, 256);
    if (msplit > max_split) msplit = max_split;
    if (msplit < 1) msplit = 1;
    int rows = vv_cdiv(M, msplit);
    rows = (rows + WK - 1) / WK * WK;
    msplit = vv_cdiv(M, rows);
    dim3 grid(nqt * nnt, msplit, p.G);
    vv_launch(k_wgrad_simt, dim3(grid), dim3(256), 0, st, p, nqt, rows);
    VV_CKL();
    return 0;
}
