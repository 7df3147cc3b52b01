// Weight gradient on tensor cores, tap-reuse tiles:   dW[t][n][k] += sum_m Gd[m, n] * A[shift(m, t), k]     (VvWGrad, common.h)
//
// Loading one shifted activation slab per (tap, 32 input channels) means 9 loads of every activation byte for a 3x3 conv: such a
// kernel runs at the TMA / L2->smem delivery rate, not at the tensor rate.  Here a CTA owns one 32-input-channel slab and ALL taps:
//   * per 128-pixel tile it loads one activation box per distinct dx, (bh + ndy - 1) pixel rows tall, laid out
//     [row][image][x] (32 channels per pixel row, MN-major: tf32 = 128-byte rows, 128B swizzle with 32B atoms, K = 8 pixels per
//     MMA; fp16 (template F16) = 64-byte rows, 64B swizzle, K = 16 pixels per MMA -- half the MMAs and half the bytes);
//   * ONE tcgen05.mma (M = 128) covers the ndy taps that share a dx: the four 32-row blocks of the M
//     operand are the same box read at starts dy * (bn*bw*row bytes) apart -- the descriptor's leading-dimension byte
//     offset IS the dy step (block 3, and block 2 for 2x2 taps, computes rows nobody reads);
//   * the output-gradient tile (NT <= 128 channels) is loaded once per pixel tile and shared by all dx;
//   * one TMEM accumulator per dx (ndx * NT <= 384 columns), reduced over the CTA's share of pixel tiles, then added to
//     dW with coalesced fp32 reductions.
// 3x3 conv, 32 -> 32 channels at 32x32: 88 KB of TMA traffic per 128 pixels instead of 240 KB.
#include "tc_common.cuh"

namespace {

struct Wg2Params {
    int B, H, W, G;
    int bw, bh, bn;                 // pixel box of one tile: bw * bh * bn == 128
    int tiles_x, tiles_y, tiles_n, m_tiles, tiles_per_split;
    int kchunks, n_tiles;
    int ndx, ndy, dy0;
    int dx[3];
    int tap[3][3];
    int N, Kt, cq;                  // cq: channels per space-to-depth phase of Gd (g_s2d), else 0
    int a_bytes, row_shift, a_stages, g_stages;
    float *dW;
    long long dw_gs;
};

constexpr int WG2_SMEM_MAX = 227 * 1024;

// MN-major operand descriptor.  Canonical layouts (in 16-byte units, cute/atom/mma_traits_sm100.hpp):
//   tf32: SWIZZLE_128B_BASE32B  ((8,n),(4,k)):((1,LBO),(8,SBO))   rows of 128 bytes, 4-row groups 512 bytes apart   (layout type 1)
//   fp16: SWIZZLE_64B           ((4,n),(8,k)):((1,LBO),(4,SBO))   rows of  64 bytes, 8-row groups 512 bytes apart   (layout type 4)
// LBO = byte distance between consecutive 32-channel blocks of the MN dimension, SBO = 512 in both.
template <bool F16>
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(F16 ? 4 : 1) << 61);
}
// instruction descriptor: D fp32, A/B tf32 or fp16, both MN-major, M = 128, N = n
template <bool F16>
__device__ __forceinline__ uint32_t idesc_mnmn(int n) {
    return (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int NT, bool F16>
__global__ void __launch_bounds__(128, 1) k_wgrad_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                                                      const Wg2Params p) {
    constexpr int ROWB = F16 ? KS * 2 : KS * 4;             // bytes of one pixel row of a 32-channel slab
    constexpr int KPIX = F16 ? 16 : 8;                      // pixels (K) per MMA: 1024 bytes of either operand
    constexpr int G_SLAB = BM * ROWB;                       // 128 pixels x 32 channels
    constexpr int NS = NT / 32;
    constexpr int G_STAGE = NS * G_SLAB;
    constexpr int TMEM_COLS = (3 * NT <= 128) ? 128 : (3 * NT <= 256 ? 256 : 512);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + vv_smem_pad(smem_raw, 1024);
    uint8_t *g_ring = smem;                                       // [g_stages][NS][128 px][32 ch]
    uint8_t *a_ring = smem + p.g_stages * G_STAGE;                // [a_stages][rows][image][x][32 ch]  (+ one row_shift of slack)
    uint8_t *tail = a_ring + p.a_stages * p.a_bytes + 2 * p.row_shift;   // slack: M blocks 2/3 of the last stage read past its box
    uint64_t *a_full = (uint64_t *)tail, *a_empty = a_full + 8, *g_full = a_empty + 8, *g_empty = g_full + 2, *accum = g_empty + 2;
    uint32_t *tmem_slot = (uint32_t *)(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int kc = blockIdx.x % p.kchunks, nt = blockIdx.x / p.kchunks;
    const int n0 = nt * NT;
    const int pt_begin = blockIdx.y * p.tiles_per_split;
    const int pt_end = min(p.m_tiles, pt_begin + p.tiles_per_split);
    const int ntiles = pt_end - pt_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.a_stages; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < p.g_stages; s++) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    vv_pdl_wait();                                           // set-up above overlaps the previous kernel's tail; global memory from here on

    if (ntiles > 0) {
        if (warp == 0) {
            if (lane == 0) {
                // ---------------- TMA producer
                int ia = 0;
                for (int ti = 0; ti < ntiles; ti++) {
                    int r = pt_begin + ti;
                    const int tx = r % p.tiles_x; r /= p.tiles_x;
                    const int ty = r % p.tiles_y; r /= p.tiles_y;
                    const int img0 = r * p.bn, y0 = ty * p.bh, x0 = tx * p.bw;
                    {   // output-gradient tile, shared by every dx of this pixel tile
                        const int s = ti % p.g_stages, round = ti / p.g_stages;
                        if (round > 0) mbar_wait(&g_empty[s], (round - 1) & 1);
                        mbar_expect_tx(&g_full[s], G_STAGE);
#pragma unroll
                        for (int j = 0; j < NS; j++) {
                            int c = n0 + j * KS, xx = x0, yy = y0;
                            if (p.cq) {
                                const int ph = c / p.cq;
                                c -= ph * p.cq;
                                xx = 2 * xx + (ph & 1);
                                yy = 2 * yy + (ph >> 1);
                            }
                            tma_load_5d(g_ring + s * G_STAGE + j * G_SLAB, &tmG, &g_full[s], c, xx, img0, yy, g);
                        }
                    }
                    for (int dxi = 0; dxi < p.ndx; dxi++, ia++) {
                        const int s = ia % p.a_stages, round = ia / p.a_stages;
                        if (round > 0) mbar_wait(&a_empty[s], (round - 1) & 1);
                        mbar_expect_tx(&a_full[s], p.a_bytes);
                        tma_load_5d(a_ring + s * p.a_bytes, &tmA, &a_full[s], kc * KS, x0 + p.dx[dxi], img0, y0 + p.dy0, g);
                    }
                }
            }
        } else if (warp == 1) {
            // ---------------- MMA issuer (warp-uniform loop, one elected lane issues)
            const uint32_t idesc = idesc_mnmn<F16>(NT);
            const uint32_t a_base = smem_u32(a_ring), g_base = smem_u32(g_ring);
            int ia = 0;
            for (int ti = 0; ti < ntiles; ti++) {
                const int sg = ti % p.g_stages;
                mbar_wait(&g_full[sg], (ti / p.g_stages) & 1);
                const uint64_t dg = desc_mn<F16>(g_base + sg * G_STAGE, G_SLAB);
                for (int dxi = 0; dxi < p.ndx; dxi++, ia++) {
                    const int s = ia % p.a_stages;
                    mbar_wait(&a_full[s], (ia / p.a_stages) & 1);
                    tc_fence_after();
                    // M operand: four 32-channel blocks = the same box, dy * row_shift bytes apart (LBO = one dy step)
                    const uint64_t da = desc_mn<F16>(a_base + s * p.a_bytes, p.row_shift);
                    const uint32_t d_tmem = tmem + dxi * NT;
#pragma unroll
                    for (int k = 0; k < BM / KPIX; k++)   // KPIX pixels = 1024 bytes per MMA in both operands
                        if (elect_one()) {
                            if (F16) tc_mma_f16(d_tmem, da + 64 * k, dg + 64 * k, idesc, (ti | k) ? 1u : 0u);
                            else tc_mma_tf32(d_tmem, da + 64 * k, dg + 64 * k, idesc, (ti | k) ? 1u : 0u);
                        }
                    if (elect_one()) tc_commit(&a_empty[s]);
                    __syncwarp();
                }
                if (elect_one()) tc_commit(&g_empty[sg]);
                __syncwarp();
            }
            if (elect_one()) tc_commit(accum);
            __syncwarp();
        }
        __syncwarp();
        // ---------------- epilogue: warp = dy, lane = input channel within the slab (contiguous in dW), column = output channel
        mbar_wait(accum, 0);
        tc_fence_after();
        if (warp < p.ndy) {
            for (int dxi = 0; dxi < p.ndx; dxi++) {
                const int t = p.tap[warp][dxi];
                float *dst = p.dW + g * p.dw_gs + ((long long)t * p.N + n0) * p.Kt + kc * KS + lane;
#pragma unroll 1
                for (int c0 = 0; c0 < NT; c0 += 32) {
                    float v[32];
                    tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + dxi * NT + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j++) atomicAdd(dst + (long long)(c0 + j) * p.Kt, v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

bool analyse_taps_wg(const VvTaps &t, Wg2Params &wp) {
    int dxs[9], ndx = 0, dymin = 99, dymax = -99;
    for (int k = 0; k < t.n; k++) {
        bool seen = false;
        for (int j = 0; j < ndx; j++) seen = seen || dxs[j] == t.dx[k];
        if (!seen) dxs[ndx++] = t.dx[k];
        dymin = t.dy[k] < dymin ? t.dy[k] : dymin;
        dymax = t.dy[k] > dymax ? t.dy[k] : dymax;
    }
    const int ndy = dymax - dymin + 1;
    if (ndx > 3 || ndy > 3 || ndx * ndy != t.n) return false;
    for (int i = 0; i < ndx; i++)
        for (int j = i + 1; j < ndx; j++)
            if (dxs[j] < dxs[i]) { int tmp = dxs[i]; dxs[i] = dxs[j]; dxs[j] = tmp; }
    wp.ndx = ndx; wp.ndy = ndy; wp.dy0 = dymin;
    for (int i = 0; i < ndx; i++) wp.dx[i] = dxs[i];
    for (int dyi = 0; dyi < ndy; dyi++)
        for (int dxi = 0; dxi < ndx; dxi++) {
            int found = -1;
            for (int k = 0; k < t.n; k++)
                if (t.dy[k] == dymin + dyi && t.dx[k] == dxs[dxi]) found = k;
            if (found < 0) return false;
            wp.tap[dyi][dxi] = found;
        }
    return true;
}

template <int NT, bool F16>
int launch_wg2(const CUtensorMap &tmA, const CUtensorMap &tmG, const Wg2Params &wp, dim3 grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_wgrad_tc2<NT, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM_MAX));
        attr = true;
    }
    vv_launch(k_wgrad_tc2<NT, F16>, dim3(grid), dim3(128), smem, st, tmA, tmG, wp);
    VV_CKL();
    return 0;
}

}  // namespace

bool vv_wgrad_tc2_supported(const VvWGrad &p) {
    static int off = -1;
    if (off < 0) {
        const char *e = getenv("VECVAD_NO_TC2");
        off = (e && e[0] == '1') ? 1 : 0;
    }
    if (off || !vv_wgrad_tc_supported(p)) return false;
    Wg2Params wp;
    int bw, bh, bn;
    if (!tile_geometry_n(p.H, p.W, BM, bw, bh, bn)) return false;
    if (bn * bw * KS * (p.ab_f16 ? 2 : 4) < 512) return false;   // a dy step must be a multiple of the 512-byte swizzle pattern
    return analyse_taps_wg(p.taps, wp);
}

int vv_launch_wgrad_tc2(const VvWGrad &p, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    Wg2Params wp;
    memset(&wp, 0, sizeof(wp));
    VV_REQUIRE(enc && analyse_taps_wg(p.taps, wp), "wgrad_tc2: unsupported tap pattern");
    wp.B = p.B; wp.H = p.H; wp.W = p.W; wp.G = p.G;
    VV_REQUIRE(tile_geometry_n(p.H, p.W, BM, wp.bw, wp.bh, wp.bn), "wgrad_tc2: unsupported image size %dx%d", p.H, p.W);
    wp.tiles_x = p.W / wp.bw; wp.tiles_y = p.H / wp.bh; wp.tiles_n = (p.B + wp.bn - 1) / wp.bn;
    wp.m_tiles = wp.tiles_x * wp.tiles_y * wp.tiles_n;
    wp.kchunks = p.Kt / KS;
    wp.N = p.N; wp.Kt = p.Kt; wp.cq = p.g_s2d ? p.N / 4 : 0;
    wp.dW = p.dW; wp.dw_gs = p.dw_gs;
    const int rows = wp.bh + wp.ndy - 1;
    const int esz = p.ab_f16 ? 2 : 4;
    wp.row_shift = wp.bn * wp.bw * KS * esz;
    wp.a_bytes = rows * wp.row_shift;
    const int nt_tile = p.N % 128 == 0 ? 128 : (p.N % 64 == 0 ? 64 : 32);
    wp.n_tiles = p.N / nt_tile;
    const int g_stage = (nt_tile / 32) * BM * KS * esz;
    const int fixed = 1024 + 256 + 2 * wp.row_shift;
    wp.g_stages = 2;
    int a_stages = (WG2_SMEM_MAX - fixed - wp.g_stages * g_stage) / wp.a_bytes;
    if (a_stages > 8) a_stages = 8;
    {
        static int cap = -1;
        if (cap < 0) { const char *e = getenv("VECVAD_WG2_STAGES"); cap = e ? atoi(e) : 0; }
        if (cap >= 2 && a_stages > cap) a_stages = cap;
        else if (cap == 0 && a_stages > 3) a_stages = 3;  // see igemm_tc2.cu: leaves shared memory for co-resident BN kernels
    }
    VV_REQUIRE(a_stages >= 2, "wgrad_tc2: tile does not fit in shared memory");
    wp.a_stages = a_stages;
    const int smem = fixed + wp.g_stages * g_stage + a_stages * wp.a_bytes;

    const CUtensorMapDataType dt = p.ab_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : tmap_dtype();
    const CUtensorMapSwizzle sw = p.ab_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    const unsigned long long e = esz;
    alignas(64) CUtensorMap tmA, tmG;
    {
        // dimensions (channel, x, image, y, group): boxes land in shared memory as [row][image][x][32 ch]; images past the
        // batch (ragged last tile) are out of bounds in their own dimension and arrive as zeros
        cuuint64_t dims[5] = {(cuuint64_t)p.Kt, (cuuint64_t)p.W, (cuuint64_t)p.B, (cuuint64_t)p.H, (cuuint64_t)p.G};
        cuuint64_t strides[4] = {(cuuint64_t)p.lda * e, (cuuint64_t)p.H * p.W * p.lda * e, (cuuint64_t)p.W * p.lda * e,
                                 (cuuint64_t)(p.G > 1 ? p.a_gs : (long long)p.B * p.H * p.W * p.lda) * e};
        cuuint32_t box[5] = {KS, (cuuint32_t)wp.bw, (cuuint32_t)wp.bn, (cuuint32_t)rows, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmA, dt, 5, (void *)((const char *)p.A + (long long)p.a_coff * esz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "wgrad_tc2: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        const int sc = p.g_s2d ? 2 : 1;
        const cuuint64_t C = p.g_s2d ? p.N / 4 : p.N;
        cuuint64_t dims[5] = {C, (cuuint64_t)sc * p.W, (cuuint64_t)p.B, (cuuint64_t)sc * p.H, (cuuint64_t)p.G};
        cuuint64_t strides[4] = {(cuuint64_t)p.ldg * e, (cuuint64_t)sc * p.H * sc * p.W * p.ldg * e, (cuuint64_t)sc * p.W * p.ldg * e,
                                 (cuuint64_t)(p.G > 1 ? p.g_gs : (long long)p.B * sc * p.H * sc * p.W * p.ldg) * e};
        cuuint32_t box[5] = {KS, (cuuint32_t)(sc * wp.bw), (cuuint32_t)wp.bn, (cuuint32_t)(sc * wp.bh), 1};
        cuuint32_t estr[5] = {1, (cuuint32_t)sc, 1, (cuuint32_t)sc, 1};
        CUresult r = enc(&tmG, dt, 5, (void *)((const char *)p.Gd + (long long)p.g_coff * esz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "wgrad_tc2: cuTensorMapEncodeTiled(Gd) failed with %d", (int)r);
    }
    const int out_tiles = wp.kchunks * wp.n_tiles * p.G;
    // CTAs per launch aimed at: two waves of shorter CTAs give SMs back to the main stream's tile kernels sooner than one wave of
    // long ones (measured, batch 128: 148 -> 2.460 ms/step, 222 -> 2.449, 296 -> 2.443, 444 -> 2.495: more splits, more dW atomics)
    static int target = -1;
    if (target < 0) { const char *e = getenv("VECVAD_WG_TARGET"); target = e ? atoi(e) : 296; if (target < 1) target = 296; }
    int splits = (target + out_tiles - 1) / out_tiles;
    if (splits > wp.m_tiles) splits = wp.m_tiles;
    if (splits < 1) splits = 1;
    wp.tiles_per_split = (wp.m_tiles + splits - 1) / splits;
    splits = (wp.m_tiles + wp.tiles_per_split - 1) / wp.tiles_per_split;
    dim3 grid(wp.kchunks * wp.n_tiles, splits, p.G);
    if (p.ab_f16) {
        if (nt_tile == 128) return launch_wg2<128, true>(tmA, tmG, wp, grid, smem, st);
        if (nt_tile == 64) return launch_wg2<64, true>(tmA, tmG, wp, grid, smem, st);
        return launch_wg2<32, true>(tmA, tmG, wp, grid, smem, st);
    }
    if (nt_tile == 128) return launch_wg2<128, false>(tmA, tmG, wp, grid, smem, st);
    if (nt_tile == 64) return launch_wg2<64, false>(tmA, tmG, wp, grid, smem, st);
    return launch_wg2<32, false>(tmA, tmG, wp, grid, smem, st);
}
