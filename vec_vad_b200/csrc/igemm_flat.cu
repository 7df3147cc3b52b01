// tcgen05 (kind::tf32) 3x3 conv tiles over a FLATTENED, zero-separated pixel sequence: ONE activation box serves all 9 taps.
//
//   out[m, n] = bias[n] + sum_t sum_k A[shift(m, t), k] * Wt[t][n][k]          (VvIGemm, common.h; 3x3 taps of either sign)
//
// k_igemm_tc2 (igemm_tc2.cu) loads one box per distinct dx because a dx step inside a [rows][W] box wraps around the image
// edge.  At 32x32 resolution that is 3 boxes of 6 rows per 4 output rows = 4.5 x the activation bytes through L2 -> shared
// memory, and shared-memory bandwidth (TMA writes + the UMMA re-reading its 128-row A slice per tap) is what bounds those
// layers.  Here the box is (W + 1) pixels wide: TMA zero-fills the out-of-range column x = W, so in
// shared memory image rows sit P = W + 1 pixel-rows apart with ONE zero pixel between them, which is both the right padding
// of row y and the left padding of row y + 1.  In that flattened sequence every tap (dy, dx) of every pixel is the constant
// offset dy * P + dx, so the A operand of a tap is the SAME box read through a UMMA descriptor that starts (dy * P + dx) * 128
// bytes further in -- 9 taps from one load.  A tile is 128 consecutive sequence positions of one image (positions on the
// separator column or past the image are computed and discarded: W/P * L/(128*ceil(L/128)) = 86 % useful rows at 32x32);
// its box is the 7 image rows those positions and their halo touch: 1.8 x the activation bytes instead of 4.5 x.
//   Descriptor starts are 128-byte but not 1024-byte aligned: the swizzle phase of the start row goes into the descriptor's
//   base-offset field (VECVAD_FLAT = 1 or 2 selects the convention; validated on the device by tests/test_conv_gpu.py).
//   warp 0: TMA producer | warp 1: MMA issuer (+ TMEM alloc) | warps 2-5: epilogue (bias, NHWC stores, BatchNorm statistics).
#include "tc_common.cuh"

#ifndef VECVAD_FLAT_DEFAULT
#define VECVAD_FLAT_DEFAULT 0
#endif

namespace {

struct FlatParams {
    int B, H, W, G;
    int P, L, tpi, m_tiles;         // row pitch W+1, positions per image H*P, tiles per image, tiles in total
    int rows;                       // image rows per box
    int kchunks, stages;
    int tap[3][3];                  // [dy+1][dx+1] -> tap index of the weight tensor
    int a_bytes, stage_bytes;
    int bo_mode, dbg;
    int N;
    float *O;
    long long o_gs;
    int ldo, o_coff;
    const float *bias;
    long long bias_gs;
    double *stats;
    long long stats_gs;
};

constexpr int FL_THREADS = 192;
constexpr int FL_SMEM_MAX = 227 * 1024;
constexpr int FL_GUARD = 1024;      // zeroed bytes in front of stage 0 (a tile starting at x = 0 reads one pixel-row before its box)

__device__ __forceinline__ uint64_t smem_desc_flat(uint32_t saddr, int bo_mode) {
    uint64_t d = smem_desc_k_sw128(saddr);
    if (bo_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;     // matrix base offset: phase of the start row in the 8-row swizzle atom
    return d;
}

template <int BN>
__global__ void __launch_bounds__(FL_THREADS, 1) k_igemm_flat(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                              const FlatParams p) {
    constexpr int B_TAP = BN * KS * 4;                       // one (tap, slab) weight tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *stage0 = smem + FL_GUARD;
    uint8_t *b_stat = stage0 + p.stages * p.stage_bytes;
    uint8_t *tail = b_stat + 9 * p.kchunks * B_TAP;
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
    for (int i = threadIdx.x; i < BN; i += FL_THREADS) {
        s_bias[i] = p.bias ? p.bias[g * p.bias_gs + n0 + i] : 0.f;
        s_sum[i] = 0.f; s_sq[i] = 0.f;
    }
    // zero what TMA never writes but the MMAs may read: the guard in front of stage 0 and the slack behind every box
    for (int i = threadIdx.x; i < FL_GUARD / 16; i += FL_THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    {
        const int slack16 = (p.stage_bytes - p.a_bytes) / 16;
        for (int i = threadIdx.x; i < p.stages * slack16; i += FL_THREADS) {
            const int s = i / slack16, j = i - s * slack16;
            reinterpret_cast<uint4 *>(stage0 + s * p.stage_bytes + p.a_bytes)[j] = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy zeros -> visible to the UMMA (async proxy) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: weights once, then one box per (tile, 32-channel slab)
            mbar_expect_tx(bfull, 9 * p.kchunks * B_TAP);
            for (int t = 0; t < 9; t++)
                for (int kc = 0; kc < p.kchunks; kc++)
                    tma_load_3d(b_stat + (t * p.kchunks + kc) * B_TAP, &tmB, bfull, kc * KS, n0, g * 9 + t);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
                const int img = tile / p.tpi, tt = tile - img * p.tpi;
                const int r_lo = (tt * BM) / p.P;                      // first image row with a position in this tile
                for (int kc = 0; kc < p.kchunks; kc++, it++) {
                    const int s = it % p.stages, round = it / p.stages;
                    if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                    mbar_expect_tx(&full[s], p.a_bytes);
                    tma_load_4d(stage0 + s * p.stage_bytes, &tmA, &full[s], kc * KS, 0, r_lo - 1, g * p.B + img);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: warp-uniform loop, one elected lane issues
        const uint32_t idesc = idesc_tf32(BN);
        const uint32_t st0 = smem_u32(stage0), bstat_base = smem_u32(b_stat);
        mbar_wait(bfull, 0);
        int it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, tcount++) {
            const int buf = tcount & 1, use = tcount >> 1;
            const int img = tile / p.tpi, tt = tile - img * p.tpi;
            const int r_lo = (tt * BM) / p.P;
            const int q0 = tt * BM - r_lo * p.P + p.P;                 // box row of the tile's first position (box row 0 = image row r_lo-1, x=0)
            if (use > 0) mbar_wait(&acc_empty[buf], (use - 1) & 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem + buf * BN;
            uint32_t acc = 0;
            for (int kc = 0; kc < p.kchunks; kc++, it++) {
                const int s = it % p.stages, round = it / p.stages;
                mbar_wait(&full[s], round & 1);
                tc_fence_after();
                const uint32_t sa = st0 + s * p.stage_bytes;
#pragma unroll
                for (int dyi = 0; dyi < 3; dyi++) {
#pragma unroll
                    for (int dxi = 0; dxi < 3; dxi++) {
                        const int rowoff = q0 + (dyi - 1) * p.P + (dxi - 1);       // >= -1: the guard / the previous stage's slack
                        const uint64_t da = smem_desc_flat(sa + rowoff * (KS * 4), p.bo_mode);
                        const uint64_t db = smem_desc_k_sw128(bstat_base + (p.tap[dyi][dxi] * p.kchunks + kc) * B_TAP);
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
            if (elect_one()) tc_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else {
        // ---------------- epilogue warps 2..5; warp w may touch TMEM lanes 32*(w%4) .. +31 (= tile rows)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float *O = p.O + g * p.o_gs;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, tcount++) {
            const int buf = tcount & 1, use = tcount >> 1;
            const int img = tile / p.tpi, tt = tile - img * p.tpi;
            const int pos = tt * BM + row;
            const int y = pos / p.P, x = pos - y * p.P;
            const bool valid = x < p.W && y < p.H;                    // not the separator column, not past the image
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * BN + c0, v);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += s_bias[c0 + j];
                if (valid && !(p.dbg & 1)) {
                    float *dst = O + ((long long)(img * p.H + y) * p.W + x) * p.ldo + p.o_coff + n0 + c0;
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

// all nine (dy, dx) in {-1,0,1}^2, in any order / sign convention
bool analyse_3x3(const VvTaps &t, FlatParams &fp) {
    if (t.n != 9) return false;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) fp.tap[i][j] = -1;
    for (int k = 0; k < 9; k++) {
        if (t.dy[k] < -1 || t.dy[k] > 1 || t.dx[k] < -1 || t.dx[k] > 1) return false;
        fp.tap[t.dy[k] + 1][t.dx[k] + 1] = k;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (fp.tap[i][j] < 0) return false;
    return true;
}

inline int flat_bn_tile(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32); }

// 0: off; 1: descriptors carry the base offset (documented convention); 2: base offset left at 0
int flat_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECVAD_FLAT");
        v = e ? atoi(e) : VECVAD_FLAT_DEFAULT;
    }
    return v;
}

template <int BN>
int launch_flat(const CUtensorMap &tmA, const CUtensorMap &tmB, const FlatParams &fp, dim3 grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_igemm_flat<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM_MAX));
        attr = true;
    }
    k_igemm_flat<BN><<<grid, FL_THREADS, smem, st>>>(tmA, tmB, fp);
    VV_CKL();
    return 0;
}

}  // namespace

// shapes the flattened-sequence tiles take: 3x3 taps, plain NHWC in/out, all nine weight tiles resident (<= 72 KB)
bool vv_igemm_flat_shape_ok(const VvIGemm &p) {
    FlatParams fp;
    if (!vv_igemm_tc_supported(p) || p.a_s2d || p.o_d2s || !analyse_3x3(p.taps, fp)) return false;
    if (p.W + 1 > 256 || p.W < 8) return false;
    const int b_all = 9 * (p.Kt / KS) * flat_bn_tile(p.N) * KS * 4;
    return b_all <= 72 * 1024;
}

// used by the net engine: only where the flattened tiles beat the per-dx boxes (full-width rows, i.e. W >= 32)
bool vv_igemm_flat_supported(const VvIGemm &p) { return flat_mode() != 0 && p.W >= 32 && vv_igemm_flat_shape_ok(p); }

int vv_launch_igemm_flat(const VvIGemm &p, cudaStream_t st, int bo_override) {
    EncodeTiledFn enc = encode_fn();
    FlatParams fp;
    memset(&fp, 0, sizeof(fp));
    VV_REQUIRE(enc && vv_igemm_flat_shape_ok(p) && analyse_3x3(p.taps, fp), "igemm_flat: unsupported shape (Kt=%d N=%d H=%d W=%d)", p.Kt, p.N,
               p.H, p.W);
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("VECVAD_DBG_TC2"); dbg = e ? atoi(e) : 0; }
        fp.dbg = dbg;
    }
    const int mode = bo_override ? bo_override : (flat_mode() ? flat_mode() : 1);
    fp.bo_mode = mode == 1 ? 1 : 0;
    fp.B = p.B; fp.H = p.H; fp.W = p.W; fp.G = p.G;
    fp.P = p.W + 1; fp.L = p.H * fp.P; fp.tpi = (fp.L + BM - 1) / BM; fp.m_tiles = fp.tpi * p.B;
    fp.rows = (BM - 1) / fp.P + 4;                    // rows holding 128 consecutive positions (<= 127/P + 2) + one halo row either side
    fp.kchunks = p.Kt / KS;
    fp.a_bytes = fp.rows * fp.P * KS * 4;
    fp.stage_bytes = (fp.a_bytes + KS * 4 + 1023) / 1024 * 1024;      // >= one zero pixel-row of slack: the next stage's leading guard
    fp.N = p.N; fp.O = p.O; fp.o_gs = p.o_gs; fp.ldo = p.ldo; fp.o_coff = p.o_coff;
    fp.bias = p.bias; fp.bias_gs = p.bias_gs; fp.stats = p.stats; fp.stats_gs = p.stats_gs;
    const int bn_tile = flat_bn_tile(p.N);
    const int b_all = 9 * fp.kchunks * bn_tile * KS * 4;
    const int fixed = 1024 /*alignment*/ + FL_GUARD + 256 /*barriers*/ + 3 * bn_tile * 4;
    int stages = (FL_SMEM_MAX - fixed - b_all) / fp.stage_bytes;
    {
        static int cap = -1;
        if (cap < 0) { const char *e = getenv("VECVAD_FLAT_STAGES"); cap = e ? atoi(e) : 3; }
        if (cap >= 2 && stages > cap) stages = cap;
    }
    if (stages > 8) stages = 8;
    VV_REQUIRE(stages >= 2, "igemm_flat: tile does not fit in shared memory");
    fp.stages = stages;
    const int smem = fixed + b_all + stages * fp.stage_bytes;

    const CUtensorMapDataType dt = tmap_dtype();
    alignas(64) CUtensorMap tmA, tmB;
    {
        // (channel, x, y, image); the box is P = W + 1 wide and starts at x = 0: column W is out of range = the zero separator
        cuuint64_t dims[4] = {(cuuint64_t)p.Kt, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.G * p.B};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * 4, (cuuint64_t)p.W * p.lda * 4, (cuuint64_t)p.H * p.W * p.lda * 4};
        cuuint32_t box[4] = {KS, (cuuint32_t)fp.P, (cuuint32_t)fp.rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, dt, 4, (void *)(p.A + p.a_coff), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_flat: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Kt, (cuuint64_t)p.N, (cuuint64_t)9 * p.G};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kt * 4, (cuuint64_t)p.N * p.Kt * 4};
        cuuint32_t box[3] = {KS, (cuuint32_t)bn_tile, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmB, dt, 3, (void *)p.Wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_flat: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    const int n_tiles = p.N / bn_tile;
    int gx = 148 / (n_tiles * p.G);
    if (gx < 1) gx = 1;
    if (gx > fp.m_tiles) gx = fp.m_tiles;
    dim3 grid(gx, n_tiles, p.G);
    if (bn_tile == 128) return launch_flat<128>(tmA, tmB, fp, grid, smem, st);
    if (bn_tile == 64) return launch_flat<64>(tmA, tmB, fp, grid, smem, st);
    return launch_flat<32>(tmA, tmB, fp, grid, smem, st);
}
