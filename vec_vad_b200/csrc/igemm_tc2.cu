// Persistent, warp-specialised tcgen05 (kind::tf32) implicit-GEMM conv tiles with tap reuse out of shared memory.
//
//   out[m, n] = bias[n] + sum_t sum_k A[shift(m, t), k] * Wt[t][n][k]          (VvIGemm, common.h)
//
// What changes against k_igemm_tc (igemm_tc.cu), which fetched every (tap, 32-channel slab) of the activations
// separately (9 x per 3x3 conv) and was bound by L2 -> shared-memory operand traffic (profiles/r01_full_tc_v1.txt):
//   * TAP REUSE.  Per 32-channel slab only ONE box per distinct dx is loaded, (bh + ndy - 1) pixel rows tall
//     (rows outside the image zero-filled by TMA = the conv's padding).  The box is laid out [row][image][x] so
//     that one dy step is a constant, 1024-byte-aligned offset: the ndy taps sharing a dx are the SAME shared-memory
//     box read through UMMA descriptors that start dy * (bn*bw*128) bytes further in.  3x3 conv: 3 loads of 6 rows
//     instead of 9 loads of 4 rows (2x less activation traffic at 32x32, 2.4x at 16x16 / 8x8).
//   * STATIONARY WEIGHTS.  When all taps x slabs of the weight tile fit in 72 KB they are loaded once per CTA.
//   * PERSISTENT CTAs, one per SM, looping over pixel tiles; two TMEM accumulators so the epilogue of tile i
//     (4 dedicated warps: tcgen05.ld, bias, NHWC stores, BatchNorm statistics) overlaps the MMAs of tile i+1.
//   warp 0: TMA producer | warp 1: MMA issuer (+ TMEM alloc) | warps 2-5: epilogue.
#include "tc_common.cuh"

namespace {

struct Tc2Params {
    int B, H, W, G;
    int bw, bh, bn;                 // pixel box of one tile: bw * bh * bn == 128
    int tiles_x, tiles_y, tiles_n, m_tiles;
    int kchunks, cq;                // cq: channels per space-to-depth phase (a_s2d), else 0
    int ndx, ndy, dy0;
    int dx[3];
    int tap[3][3];                  // [dyi][dxi] -> tap index of the weight tensor
    int ntaps;
    int stages, stationary;
    int dbg;                        // VECVAD_DBG_TC2 bits (timing experiments only): 1 skip stores, 2 skip statistics, 4 skip MMAs
    int a_bytes, row_shift;         // bytes of one activation box; bytes of one dy step (= bn*bw*128)
    int N;
    float *O;
    long long o_gs;
    int ldo, o_coff, o_d2s;
    const float *bias;
    long long bias_gs;
    double *stats;
    long long stats_gs;
};

constexpr int T2_THREADS = 192;
constexpr int SMEM_MAX = 227 * 1024;

template <int BN>
__global__ void __launch_bounds__(T2_THREADS, 1) k_igemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                             const Tc2Params p) {
    constexpr int B_TAP = BN * KS * 4;                       // one (tap, slab) weight tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = p.a_bytes + (p.stationary ? 0 : p.ndy * B_TAP);
    uint8_t *b_stat = smem + p.stages * stage_bytes;
    uint8_t *tail = b_stat + (p.stationary ? p.ntaps * p.kchunks * B_TAP : 0);
    uint64_t *full = (uint64_t *)tail;                       // [stages]
    uint64_t *empty = full + 8;                              // [stages]
    uint64_t *acc_full = empty + 8;                          // [2]
    uint64_t *acc_empty = acc_full + 2;                      // [2]
    uint64_t *bfull = acc_empty + 2;                         // [1]
    uint32_t *tmem_slot = (uint32_t *)(bfull + 1);
    float *s_bias = (float *)(tail + 256);
    float *s_sum = s_bias + BN, *s_sq = s_sum + BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        mbar_init(bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {   // two accumulators of BN fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < BN; i += T2_THREADS) {
        s_bias[i] = p.bias ? p.bias[g * p.bias_gs + (p.o_d2s ? (n0 + i) % (p.N >> 2) : (n0 + i))] : 0.f;
        s_sum[i] = 0.f; s_sq[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            if (p.stationary) {
                mbar_expect_tx(bfull, p.ntaps * p.kchunks * B_TAP);
                for (int t = 0; t < p.ntaps; t++)
                    for (int kc = 0; kc < p.kchunks; kc++)
                        tma_load_3d(b_stat + (t * p.kchunks + kc) * B_TAP, &tmB, bfull, kc * KS, n0, g * p.ntaps + t);
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                int r = tile;
                const int tx = r % p.tiles_x; r /= p.tiles_x;
                const int ty = r % p.tiles_y; r /= p.tiles_y;
                const int img0 = r * p.bn, y0 = ty * p.bh, x0 = tx * p.bw;
                for (int kc = 0; kc < p.kchunks; kc++) {
                    for (int dxi = 0; dxi < p.ndx; dxi++, it++) {
                        const int s = it % p.stages, round = it / p.stages;
                        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                        uint8_t *sa = smem + s * stage_bytes;
                        mbar_expect_tx(&full[s], stage_bytes);
                        int c = kc * KS, xx = x0 + p.dx[dxi], yy = y0 + p.dy0;
                        if (p.cq) {                 // space-to-depth source: slab -> (phase, channel), stride-2 pixel walk
                            const int ph = c / p.cq;
                            c -= ph * p.cq;
                            xx = 2 * xx + (ph & 1);
                            yy = 2 * yy + (ph >> 1);
                        }
                        tma_load_4d(sa, &tmA, &full[s], c, xx, g * p.B + img0, yy);
                        if (!p.stationary)
                            for (int dyi = 0; dyi < p.ndy; dyi++)
                                tma_load_3d(sa + p.a_bytes + dyi * B_TAP, &tmB, &full[s], kc * KS, n0, g * p.ntaps + p.tap[dyi][dxi]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer.  The whole warp walks the loop (everything stays warp-uniform: descriptors are built in
        // uniform registers, no per-instruction lane->uniform moves); one elected lane issues the tcgen05 instructions.
        const uint32_t idesc = idesc_tf32(BN);
        const uint32_t smem_base = smem_u32(smem), bstat_base = smem_u32(b_stat);
        if (p.stationary) mbar_wait(bfull, 0);
        int it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, tcount++) {
            const int buf = tcount & 1, use = tcount >> 1;
            if (use > 0) mbar_wait(&acc_empty[buf], (use - 1) & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem + buf * BN;
            uint32_t acc = 0;
            for (int kc = 0; kc < p.kchunks; kc++) {
                for (int dxi = 0; dxi < p.ndx; dxi++, it++) {
                    const int s = it % p.stages, round = it / p.stages;
                    mbar_wait(&full[s], round & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_base + ((p.dbg & 8) ? 0 : s * stage_bytes);
#pragma unroll
                    for (int dyi = 0; dyi < 3; dyi++) {
                        if (dyi < p.ndy) {
                            const uint32_t sb = p.stationary ? bstat_base + (p.tap[dyi][dxi] * p.kchunks + kc) * B_TAP : sa + p.a_bytes + dyi * B_TAP;
                            const uint64_t da = smem_desc_k_sw128(sa + dyi * p.row_shift), db = smem_desc_k_sw128(sb);
                            if (!(p.dbg & 4)) {
#pragma unroll
                                for (int k = 0; k < KS / 8; k++) {
                                    if (elect_one()) tc_mma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (acc | k) ? 1u : 0u);
                                }
                            }
                            acc = 1;
                        }
                    }
                    if (elect_one()) tc_commit(&empty[s]);
                    __syncwarp();
                }
            }
            if (elect_one()) tc_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else {
        // ---------------- epilogue warps 2..5; warp w may touch TMEM lanes 32*(w%4) .. +31 (= tile rows)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int xx = row % p.bw, nn = (row / p.bw) % p.bn, yy = row / (p.bw * p.bn);
        float *O = p.O + g * p.o_gs;
        const int Co = p.o_d2s ? (p.N >> 2) : p.N;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, tcount++) {
            const int buf = tcount & 1, use = tcount >> 1;
            int r = tile;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y; r /= p.tiles_y;
            const int b = r * p.bn + nn, y = ty * p.bh + yy, x = tx * p.bw + xx;
            const bool valid = b < p.B;
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * BN + c0, v);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += s_bias[c0 + j];
                if (valid && !(p.dbg & 1)) {
                    float *dst;
                    const int ncol = n0 + c0;
                    if (!p.o_d2s) {
                        dst = O + ((long long)(b * p.H + y) * p.W + x) * p.ldo + p.o_coff + ncol;
                    } else {     // N = 4*Co: column block (phase, co) -> pixel (2y+py, 2x+px); a 32-column block never straddles phases
                        const int ph = ncol / Co, co = ncol - ph * Co;
                        dst = O + ((long long)(b * 2 * p.H + 2 * y + (ph >> 1)) * (2 * p.W) + 2 * x + (ph & 1)) * p.ldo + p.o_coff + co;
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                if (p.stats && !(p.dbg & 2)) {
                    // column sums over this warp's 32 rows: butterfly transpose-reduce (31 shuffles per quantity); lane j gets column c0+j
                    float s[32], sq[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) { s[j] = valid ? v[j] : 0.f; sq[j] = s[j] * s[j]; }
#pragma unroll
                    for (int w = 16; w >= 1; w >>= 1) {
                        const bool up = lane & w;
#pragma unroll
                        for (int j = 0; j < w; j++) {
                            float keep_s = up ? s[j + w] : s[j], send_s = up ? s[j] : s[j + w];
                            float keep_q = up ? sq[j + w] : sq[j], send_q = up ? sq[j] : sq[j + w];
                            s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
                            sq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
                        }
                    }
                    atomicAdd(&s_sum[c0 + lane], s[0]);
                    atomicAdd(&s_sq[c0 + lane], sq[0]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);       // this warp has drained its quarter of the accumulator
        }
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
            double *st = p.stats + g * p.stats_gs;
            for (int i = threadIdx.x - 64; i < BN; i += 128) {
                if (n0 + i < p.N) {
                    atomicAdd(&st[n0 + i], (double)s_sum[i]);
                    atomicAdd(&st[p.N + n0 + i], (double)s_sq[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
}

// taps must form a full (dy range) x (dx set) grid: true for 3x3 (both orientations) and the 2x2 phase taps of the transposed conv
bool analyse_taps(const VvTaps &t, Tc2Params &tp) {
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
    tp.ndx = ndx; tp.ndy = ndy; tp.dy0 = dymin; tp.ntaps = t.n;
    for (int i = 0; i < ndx; i++) tp.dx[i] = dxs[i];
    for (int dyi = 0; dyi < ndy; dyi++)
        for (int dxi = 0; dxi < ndx; dxi++) {
            int found = -1;
            for (int k = 0; k < t.n; k++)
                if (t.dy[k] == dymin + dyi && t.dx[k] == dxs[dxi]) found = k;
            if (found < 0) return false;
            tp.tap[dyi][dxi] = found;
        }
    return true;
}

bool g_tc2_disabled = false;      // set when the permuted-dimension tensor map is refused by the driver

template <int BN>
int launch2(const CUtensorMap &tmA, const CUtensorMap &tmB, const Tc2Params &tp, dim3 grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_igemm_tc2<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        attr = true;
    }
    k_igemm_tc2<BN><<<grid, T2_THREADS, smem, st>>>(tmA, tmB, tp);
    VV_CKL();
    return 0;
}

}  // namespace

bool vv_igemm_tc2_supported(const VvIGemm &p) {
    static int off = -1;
    if (off < 0) {
        const char *e = getenv("VECVAD_NO_TC2");
        off = (e && e[0] == '1') ? 1 : 0;
    }
    if (off || g_tc2_disabled || !vv_igemm_tc_supported(p)) return false;
    Tc2Params tp;
    return analyse_taps(p.taps, tp);
}

int vv_launch_igemm_tc2(const VvIGemm &p, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    Tc2Params tp;
    memset(&tp, 0, sizeof(tp));
    VV_REQUIRE(enc && analyse_taps(p.taps, tp), "igemm_tc2: unsupported tap pattern");
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("VECVAD_DBG_TC2"); dbg = e ? atoi(e) : 0; }
        tp.dbg = dbg;
    }
    tp.B = p.B; tp.H = p.H; tp.W = p.W; tp.G = p.G;
    VV_REQUIRE(tile_geometry_n(p.H, p.W, BM, tp.bw, tp.bh, tp.bn), "igemm_tc2: unsupported image size %dx%d", p.H, p.W);
    tp.tiles_x = p.W / tp.bw; tp.tiles_y = p.H / tp.bh; tp.tiles_n = (p.B + tp.bn - 1) / tp.bn;
    tp.m_tiles = tp.tiles_x * tp.tiles_y * tp.tiles_n;
    tp.kchunks = p.Kt / KS; tp.cq = p.a_s2d ? p.Kt / 4 : 0;
    tp.N = p.N; tp.O = p.O; tp.o_gs = p.o_gs; tp.ldo = p.ldo; tp.o_coff = p.o_coff; tp.o_d2s = p.o_d2s;
    tp.bias = p.bias; tp.bias_gs = p.bias_gs; tp.stats = p.stats; tp.stats_gs = p.stats_gs;
    const int rows = tp.bh + tp.ndy - 1;
    tp.row_shift = tp.bn * tp.bw * KS * 4;
    tp.a_bytes = rows * tp.row_shift;
    const int bn_tile = p.N % 128 == 0 ? 128 : (p.N % 64 == 0 ? 64 : 32);
    const int b_tap = bn_tile * KS * 4;
    const int b_all = tp.ntaps * tp.kchunks * b_tap;
    const int fixed = 1024 /*alignment*/ + 256 /*barriers*/ + 3 * bn_tile * 4;
    tp.stationary = b_all <= 72 * 1024;
    const int stage_bytes = tp.a_bytes + (tp.stationary ? 0 : tp.ndy * b_tap);
    int stages = (SMEM_MAX - fixed - (tp.stationary ? b_all : 0)) / stage_bytes;
    if (stages > 8) stages = 8;
    {
        static int cap = -1;
        if (cap < 0) { const char *e = getenv("VECVAD_TC2_STAGES"); cap = e ? atoi(e) : 0; }
        if (cap >= 2 && stages > cap) stages = cap;
        else if (cap == 0 && stages > 3) stages = 3;      // deeper rings buy nothing (measured) and their shared memory keeps
                                                          // the elementwise kernels of the other streams off the SM
    }
    VV_REQUIRE(stages >= 2, "igemm_tc2: tile does not fit in shared memory");
    tp.stages = stages;
    const int smem = fixed + (tp.stationary ? b_all : 0) + stages * stage_bytes;

    const CUtensorMapDataType dt = tmap_dtype();
    alignas(64) CUtensorMap tmA, tmB;
    {
        // dimensions ordered (channel, x, image, y): the box lands in shared memory as [row][image][x][32 ch]
        const int sc = p.a_s2d ? 2 : 1;
        const cuuint64_t C = p.a_s2d ? p.Kt / 4 : p.Kt;
        cuuint64_t dims[4] = {C, (cuuint64_t)sc * p.W, (cuuint64_t)p.G * p.B, (cuuint64_t)sc * p.H};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * 4, (cuuint64_t)sc * p.H * sc * p.W * p.lda * 4, (cuuint64_t)sc * p.W * p.lda * 4};
        cuuint32_t box[4] = {KS, (cuuint32_t)(sc * tp.bw), (cuuint32_t)tp.bn, (cuuint32_t)(sc * rows)};
        cuuint32_t estr[4] = {1, (cuuint32_t)sc, 1, (cuuint32_t)sc};
        CUresult r = enc(&tmA, dt, 4, (void *)(p.A + p.a_coff), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            g_tc2_disabled = true;                // fall back to the per-tap kernel from now on
            return vv_launch_igemm_tc(p, st);
        }
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Kt, (cuuint64_t)p.N, (cuuint64_t)p.taps.n * p.G};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kt * 4, (cuuint64_t)p.N * p.Kt * 4};
        cuuint32_t box[3] = {KS, (cuuint32_t)bn_tile, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmB, dt, 3, (void *)p.Wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_tc2: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    const int n_tiles = p.N / bn_tile;
    int gx = 148 / (n_tiles * p.G);
    {
        static int div = -1;
        if (div < 0) { const char *e = getenv("VECVAD_TC2_GRID_DIV"); div = e ? atoi(e) : 1; }
        if (div > 1) gx /= div;
    }
    if (gx < 1) gx = 1;
    if (gx > tp.m_tiles) gx = tp.m_tiles;
    dim3 grid(gx, n_tiles, p.G);
    if (bn_tile == 128) return launch2<128>(tmA, tmB, tp, grid, smem, st);
    if (bn_tile == 64) return launch2<64>(tmA, tmB, tp, grid, smem, st);
    return launch2<32>(tmA, tmB, tp, grid, smem, st);
}
