// Persistent tcgen05 (kind::tf32 / kind::f16) implicit-GEMM conv tiles: a CTA works on a PAIR of 128-pixel tiles at a time
// that share every weight tile.
//
//   out[m, n] = bias[n] + sum_t sum_k A[shift(m, t), k] * Wt[t][n][k]          (VvIGemm, common.h)
//
// What the ncu pass over one train step showed for single-tile CTAs at 16x16 / 8x8 / 4x4 resolution (profiles/r01_flat_v4_step.txt):
// the launches move 6.6 - 9.3 TB/s from L2 into shared memory, i.e. they sit on the L2 -> SM delivery limit, and 55 - 70 % of
// those bytes are WEIGHTS: the 9 x K x N weight tensor does not fit next to the activation ring, so it is streamed again for
// every 128-pixel tile.  And the flattened-sequence kernel (igemm_flat.cu) showed that one warp cannot issue MMAs faster than
// one per ~75 cycles while the pipe retires them in 48 (N = 64) / 64 (N = 128), and that an epilogue sharing the SM with
// back-to-back UMMAs needs ~2300 cycles per 128 x 32 block.  Hence:
//   * PAIRS.  One ring stage = the weight tiles of one (32-channel slab, dx) + the activation boxes of BOTH tiles: half the
//     weight traffic per pixel.
//   * TWO MMA WARPS, one per tile of the pair, each with its own accumulator; both walk the same stages in the same order
//     (the stage is released when both have committed), so no consumer ever skips a barrier phase.
//   * TWO EPILOGUE SETS of four warps, one per tile of the pair: tcgen05.ld, release the accumulator at once, then bias,
//     BatchNorm statistics and NHWC stores (or the transposed conv's pixel shuffle) through an XOR-swizzled 4 KB staging tile
//     per warp: column sums without shuffles, full 128-byte lines per store instruction.
//   * Tap reuse: per slab one box per distinct dx, (bh + ndy - 1) pixel rows tall, laid out [row][image][x];
//     the dy taps are descriptor start offsets into the same box.  Weights resident when all taps x slabs fit in 72 KB.
// warp 0: TMA producer | warps 1-2: MMA issuers (warp 1 allocates TMEM) | warps 3-6 / 7-10: epilogue of tile 0 / 1 of the pair.
#include "tc_common.cuh"

namespace {

struct Tc3Params {
    int B, H, W, G;
    int bw, bh, bn;                 // pixel box of one tile: bw * bh * bn == 128
    int tiles_x, tiles_y, tiles_n, m_tiles, pairs;
    int kchunks, cq;                // cq: channels per space-to-depth phase (a_s2d), else 0
    int ndx, ndy, dy0;
    int dx[3];
    int tap[3][3];                  // [dyi][dxi] -> tap index of the weight tensor
    int ntaps;
    int stages, stationary;
    int a_bytes, row_shift, stage_bytes;   // one activation box; one dy step (= bn*bw*128); weights (if streamed) + two boxes
    unsigned long long *trace;
    int N;
    float *O;
    long long o_gs;
    int ldo, o_coff, o_d2s;
    const float *bias;
    long long bias_gs;
    double *stats;
    long long stats_gs;
    int o_f16;                      // O is fp16 (the transposed conv writing the concat half)
    int o_split;                    // columns >= o_split go to O2 as fp16 (VvIGemm::o_split); 0 = none
    __half *O2;
    long long o2_gs;
    int ldo2;
    int rev;                        // walk the pairs from the last to the first
};

constexpr int T3_THREADS = 352;
constexpr int T3_SMEM_MAX = 227 * 1024;
constexpr int T3_STG_BYTES = 32 * 32 * 4;          // per epilogue warp: 32 rows x 32 floats, 16-byte chunks XOR-swizzled by row
constexpr int T3_MAX_STAGES = 4;

// float index of 16-byte chunk c4 (0..7) of staging row r
__device__ __forceinline__ int stg_idx(int r, int c4) { return r * 32 + ((c4 ^ (r & 7)) << 2); }

// F16: operands are fp16 in HBM and shared memory (64-byte pixel rows, SWIZZLE_64B, kind::f16 with K = 16 per MMA), accumulation
// fp32; outputs fp32, or fp16 where the consumer is another contraction (o_f16 / o_split).
template <int BN, bool F16>
__global__ void __launch_bounds__(T3_THREADS, 1) k_igemm_tc3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                             const Tc3Params p) {
    constexpr int ROWB = F16 ? KS * 2 : KS * 4;              // bytes of one pixel row of a 32-channel slab
    constexpr int KSTEPS = F16 ? KS / 16 : KS / 8;           // MMAs per (tap, slab): 32 bytes of K each
    constexpr int B_TAP = BN * ROWB;                         // one (tap, slab) weight tile
    constexpr int NBUF = (4 * BN <= 512) ? 2 : 1;              // accumulator double-buffering across pairs (BN = 128: all 512 TMEM columns)
    constexpr int TMEM_COLS = 2 * NBUF * BN < 32 ? 32 : 2 * NBUF * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + vv_smem_pad(smem_raw, 1024);
    const int w_stage = p.stationary ? 0 : p.ndy * B_TAP;    // streamed weights lead each stage
    uint8_t *b_stat = smem + p.stages * p.stage_bytes;
    uint8_t *stg_base = b_stat + (p.stationary ? p.ntaps * p.kchunks * B_TAP : 0);
    uint8_t *tail = stg_base + 8 * T3_STG_BYTES;
    uint64_t *full = (uint64_t *)tail;                       // [T3_MAX_STAGES]
    uint64_t *empty = full + T3_MAX_STAGES;                  // [T3_MAX_STAGES]
    uint64_t *acc_full = empty + T3_MAX_STAGES;              // [2 buf][2 tile]
    uint64_t *acc_empty = acc_full + 4;                      // [2 buf][2 tile]
    uint64_t *bfull = acc_empty + 4;                         // [1]
    uint32_t *tmem_slot = (uint32_t *)(bfull + 1);
    float *s_bias = (float *)(tail + 256);
    float *s_sum = s_bias + BN, *s_sq = s_sum + BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    const bool tracer = p.trace && blockIdx.x + blockIdx.y + blockIdx.z == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T3_MAX_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }
        for (int b = 0; b < 4; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        mbar_init(bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    vv_pdl_wait();                                           // set-up above overlaps the previous kernel's tail; global memory from here on
    for (int i = threadIdx.x; i < BN; i += T3_THREADS) {
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
            int s = 0, round = 0;
            long long t_wait = 0;
            const long long t_begin = clock64();
            for (int lp = blockIdx.x; lp < p.pairs; lp += gridDim.x) {
                const int pair = p.rev ? p.pairs - 1 - lp : lp;
                int img0[2], y0[2], x0[2];
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    int r = 2 * pair + w;                    // a phantom second tile (odd tile count) lands past the last image: masked
                    x0[w] = (r % p.tiles_x) * p.bw; r /= p.tiles_x;
                    y0[w] = (r % p.tiles_y) * p.bh; r /= p.tiles_y;
                    img0[w] = r * p.bn;
                }
                for (int kc = 0; kc < p.kchunks; kc++) {
                    for (int dxi = 0; dxi < p.ndx; dxi++) {
                        if (round > 0) {
                            const long long t0 = clock64();
                            mbar_wait(&empty[s], (round - 1) & 1);
                            t_wait += clock64() - t0;
                        }
                        uint8_t *st = smem + s * p.stage_bytes;
                        mbar_expect_tx(&full[s], p.stage_bytes);
#pragma unroll
                        for (int w = 0; w < 2; w++) {
                            int c = kc * KS, xx = x0[w] + p.dx[dxi], yy = y0[w] + p.dy0;
                            if (p.cq) {                 // space-to-depth source: slab -> (phase, channel), stride-2 pixel walk
                                const int ph = c / p.cq;
                                c -= ph * p.cq;
                                xx = 2 * xx + (ph & 1);
                                yy = 2 * yy + (ph >> 1);
                            }
                            tma_load_4d(st + w_stage + w * p.a_bytes, &tmA, &full[s], c, xx, g * p.B + img0[w], yy);
                        }
                        if (!p.stationary)
                            for (int dyi = 0; dyi < p.ndy; dyi++)
                                tma_load_3d(st + dyi * B_TAP, &tmB, &full[s], kc * KS, n0, g * p.ntaps + p.tap[dyi][dxi]);
                        if (++s == p.stages) { s = 0; round++; }
                    }
                }
            }
            if (tracer) { p.trace[0] = t_wait; p.trace[1] = clock64() - t_begin; }
        }
    } else if (warp <= 2) {
        // ---------------- MMA issuers: warp 1 computes tile 0 of every pair, warp 2 tile 1; same stages, same order.
        const int mw = warp - 1;
        const uint32_t idesc = F16 ? idesc_f16(BN) : idesc_tf32(BN);
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t bstat_lo = (smem_u32(b_stat) & 0x3FFFF) >> 4;
        // 8-row groups 8 * ROWB bytes apart, version 1, SWIZZLE_128B (layout 2) / SWIZZLE_64B (layout 4)
        const uint32_t desc_hi = (uint32_t)((8 * ROWB) >> 4) | (1u << 14) | ((F16 ? 4u : 2u) << 29);
        long long t_wacc = 0, t_wfull = 0;
        const long long t_begin = clock64();
        if (p.stationary) mbar_wait(bfull, 0);
        int s = 0, ph = 0, pcount = 0;
        for (int lp = blockIdx.x; lp < p.pairs; lp += gridDim.x, pcount++) {
            const int buf = (NBUF == 2) ? (pcount & 1) : 0, use = (NBUF == 2) ? (pcount >> 1) : pcount;
            if (use > 0) {
                const long long t0 = clock64();
                mbar_wait(&acc_empty[buf * 2 + mw], (use - 1) & 1);
                t_wacc += clock64() - t0;
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem + (buf * 2 + mw) * BN;
            uint32_t acc = 0;
            for (int kc = 0; kc < p.kchunks; kc++) {
                for (int dxi = 0; dxi < p.ndx; dxi++) {
                    {
                        const long long t0 = clock64();
                        mbar_wait(&full[s], ph);
                        t_wfull += clock64() - t0;
                    }
                    tc_fence_after();
                    const uint32_t st_lo = ((smem_base + s * p.stage_bytes) & 0x3FFFF) >> 4;
                    const uint32_t a_lo0 = st_lo + (uint32_t)((w_stage + mw * p.a_bytes) >> 4);
#pragma unroll
                    for (int dyi = 0; dyi < 3; dyi++) {
                        if (dyi < p.ndy) {
                            const uint32_t a_lo = a_lo0 + (uint32_t)((dyi * p.row_shift) >> 4);
                            const uint32_t b_lo = p.stationary ? bstat_lo + (uint32_t)(((p.tap[dyi][dxi] * p.kchunks + kc) * B_TAP) >> 4)
                                                               : st_lo + (uint32_t)((dyi * B_TAP) >> 4);
#pragma unroll
                            for (int k = 0; k < KSTEPS; k++) {
                                const uint64_t da = ((uint64_t)desc_hi << 32) | (a_lo + 2 * k | (1u << 16));
                                const uint64_t db = ((uint64_t)desc_hi << 32) | (b_lo + 2 * k | (1u << 16));
                                if (elect_one()) {
                                    if (F16) tc_mma_f16(d_tmem, da, db, idesc, (acc | (uint32_t)k) ? 1u : 0u);
                                    else tc_mma_tf32(d_tmem, da, db, idesc, (acc | (uint32_t)k) ? 1u : 0u);
                                }
                            }
                            acc = 1;
                        }
                    }
                    if (elect_one()) tc_commit(&empty[s]);
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
            if (elect_one()) tc_commit(&acc_full[buf * 2 + mw]);
            __syncwarp();
        }
        if (tracer && lane == 0) {
            unsigned long long *tr = p.trace + 2 + 4 * mw;
            tr[0] = t_wacc; tr[1] = t_wfull; tr[2] = clock64() - t_begin; tr[3] = pcount;
        }
    } else {
        // ---------------- epilogue warps 3..6 (tile 0 of each pair) and 7..10 (tile 1); warp w may touch TMEM lanes
        // 32*(w%4) .. +31 (= tile rows)
        const int q = warp & 3, es = (warp - 3) >> 2;
        const int row = q * 32 + lane;
        const int xx = row % p.bw, nn = (row / p.bw) % p.bn, yy = row / (p.bw * p.bn);
        float *O = p.O + g * p.o_gs + p.o_coff;
        float *stg = (float *)(stg_base + (warp - 3) * T3_STG_BYTES);
        const int Co = p.o_d2s ? (p.N >> 2) : p.N;
        float c_sum[BN / 32], c_sq[BN / 32];                 // lane j: running sums of columns c0 + j over this warp's rows
#pragma unroll
        for (int c = 0; c < BN / 32; c++) { c_sum[c] = 0.f; c_sq[c] = 0.f; }
        long long t_wfull = 0;
        const long long t_begin = clock64();
        int pcount = 0;
        for (int lp = blockIdx.x; lp < p.pairs; lp += gridDim.x, pcount++) {
            const int pair = p.rev ? p.pairs - 1 - lp : lp;
            const int buf = (NBUF == 2) ? (pcount & 1) : 0, use = (NBUF == 2) ? (pcount >> 1) : pcount;
            int r = 2 * pair + es;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y; r /= p.tiles_y;
            const int b = r * p.bn + nn, y = ty * p.bh + yy, x = tx * p.bw + xx;
            const bool valid = b < p.B;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const int pix = p.o_d2s ? (b * 2 * p.H + 2 * y) * (2 * p.W) + 2 * x : (b * p.H + y) * p.W + x;
            {
                const long long t0 = clock64();
                mbar_wait(&acc_full[buf * 2 + es], use & 1);
                t_wfull += clock64() - t0;
            }
            tc_fence_after();
            float v[BN];
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) tc_ld32_nowait(tmem + ((uint32_t)(q * 32) << 16) + (buf * 2 + es) * BN + c0, v + c0);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf * 2 + es]);      // accumulator quarter is in registers: the MMA warp may reuse it
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                // bias, then this warp's 32 x 32 block into the staging tile (rows past the batch as zeros)
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(s_bias + c0 + j);
                    float4 o;
                    o.x = valid ? v[c0 + j] + b4.x : 0.f;
                    o.y = valid ? v[c0 + j + 1] + b4.y : 0.f;
                    o.z = valid ? v[c0 + j + 2] + b4.z : 0.f;
                    o.w = valid ? v[c0 + j + 3] + b4.w : 0.f;
                    *reinterpret_cast<float4 *>(stg + stg_idx(lane, j >> 2)) = o;
                }
                __syncwarp();
                if (p.stats) {                           // lane j sums column c0 + j over the 32 rows
                    float s = 0.f, sq = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; rr++) {
                        const float xv = stg[stg_idx(rr, lane >> 2) + (lane & 3)];
                        s += xv;
                        sq = fmaf(xv, xv, sq);
                    }
                    c_sum[c0 / 32] += s;
                    c_sq[c0 / 32] += sq;
                }
                // destination of this 32-column block: plain NHWC, or (phase, co) -> pixel (2y+py, 2x+px) of the transposed conv
                const int ncol = n0 + c0;
                int pixc = pix, col = ncol;
                if (p.o_d2s) {
                    const int phs = ncol / Co;
                    col = ncol - phs * Co;
                    pixc = pix + (phs >> 1) * (2 * p.W) + (phs & 1);
                }
                if (p.o_f16 || (p.o_split && ncol >= p.o_split)) {      // fp16 destination: 8 bytes per lane, 64-byte pixel rows
                    const bool side = p.o_split && ncol >= p.o_split;
                    __half *Oh = side ? p.O2 + g * p.o2_gs + (ncol - p.o_split) + 4 * (lane & 7)
                                      : reinterpret_cast<__half *>(p.O) + g * p.o_gs + p.o_coff + col + 4 * (lane & 7);
                    const int ldh = side ? p.ldo2 : p.ldo;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int rr = 4 * i + (lane >> 3);
                        const float4 o = *reinterpret_cast<const float4 *>(stg + stg_idx(rr, lane & 7));
                        const int rp = __shfl_sync(0xffffffffu, pixc, rr);
                        if ((vmask >> rr) & 1) *reinterpret_cast<uint2 *>(Oh + (long long)rp * ldh) = pack_half4(o);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) {            // store instruction i: rows 4i .. 4i+3, eight lanes per 128-byte pixel row
                        const int rr = 4 * i + (lane >> 3);
                        const float4 o = *reinterpret_cast<const float4 *>(stg + stg_idx(rr, lane & 7));
                        const int rp = __shfl_sync(0xffffffffu, pixc, rr);
                        if ((vmask >> rr) & 1) *reinterpret_cast<float4 *>(O + (long long)rp * p.ldo + col + 4 * (lane & 7)) = o;
                    }
                }
                __syncwarp();
            }
        }
        if (tracer && threadIdx.x == 96) { p.trace[10] = t_wfull; p.trace[11] = clock64() - t_begin; }
        if (p.stats) {
#pragma unroll
            for (int c = 0; c < BN / 32; c++) {
                atomicAdd(&s_sum[c * 32 + lane], c_sum[c]);
                atomicAdd(&s_sq[c * 32 + lane], c_sq[c]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight epilogue warps only
            double *st = p.stats + g * p.stats_gs;
            for (int i = threadIdx.x - 96; i < BN; i += 256) {
                if (n0 + i < p.N) {
                    atomicAdd(&st[n0 + i], (double)s_sum[i]);
                    atomicAdd(&st[p.N + n0 + i], (double)s_sq[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// taps must form a full (dy range) x (dx set) grid: true for 3x3 (both orientations) and the 2x2 phase taps of the transposed conv
bool analyse_taps3(const VvTaps &t, Tc3Params &tp) {
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

inline int tc3_bn_tile(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32); }

// fills the geometry / shared-memory plan; false when a pair stage does not fit twice
bool plan3(const VvIGemm &p, Tc3Params &tp, int &smem_bytes) {
    if (!analyse_taps3(p.taps, tp)) return false;
    const int esz = p.ab_f16 ? 2 : 4;                  // operand element size
    if (p.ab_f16 && (p.lda % 8 || p.a_coff % 8)) return false;
    if (p.o_f16 && (p.ldo % 8 || p.o_coff % 8)) return false;
    tp.B = p.B; tp.H = p.H; tp.W = p.W; tp.G = p.G;
    if (!tile_geometry_n(p.H, p.W, BM, tp.bw, tp.bh, tp.bn)) return false;
    tp.tiles_x = p.W / tp.bw; tp.tiles_y = p.H / tp.bh; tp.tiles_n = (p.B + tp.bn - 1) / tp.bn;
    tp.m_tiles = tp.tiles_x * tp.tiles_y * tp.tiles_n;
    tp.pairs = (tp.m_tiles + 1) / 2;
    tp.kchunks = p.Kt / KS; tp.cq = p.a_s2d ? p.Kt / 4 : 0;
    const int rows = tp.bh + tp.ndy - 1;
    tp.row_shift = tp.bn * tp.bw * KS * esz;
    if (tp.row_shift % 1024) return false;             // a dy step must keep the swizzle phase (and the stage layout 1024-byte aligned)
    tp.a_bytes = rows * tp.row_shift;
    const int bn_tile = tc3_bn_tile(p.N);
    const int b_tap = bn_tile * KS * esz;
    const int b_all = tp.ntaps * tp.kchunks * b_tap;
    const int fixed = 1024 /*alignment*/ + 8 * T3_STG_BYTES + 256 /*barriers*/ + 3 * bn_tile * 4;
    tp.stationary = b_all <= 72 * 1024;
    tp.stage_bytes = 2 * tp.a_bytes + (tp.stationary ? 0 : tp.ndy * b_tap);
    int stages = (T3_SMEM_MAX - fixed - (tp.stationary ? b_all : 0)) / tp.stage_bytes;
    static int cap = -1;
    if (cap < 0) { const char *e = getenv("VECVAD_TC3_STAGES"); cap = e ? atoi(e) : 3; }
    if (cap >= 2 && stages > cap) stages = cap;
    if (stages > T3_MAX_STAGES) stages = T3_MAX_STAGES;
    if (stages < 2) return false;
    tp.stages = stages;
    smem_bytes = fixed + (tp.stationary ? b_all : 0) + stages * tp.stage_bytes;
    return true;
}


template <int BN, bool F16>
int launch3(const CUtensorMap &tmA, const CUtensorMap &tmB, const Tc3Params &tp, dim3 grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_igemm_tc3<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_MAX));
        attr = true;
    }
    vv_launch(k_igemm_tc3<BN, F16>, dim3(grid), dim3(T3_THREADS), smem, st, tmA, tmB, tp);
    VV_CKL();
    if (tp.trace) {      // debugging aid: synchronous
        unsigned long long h[12];
        VV_CK(cudaStreamSynchronize(st));
        VV_CK(cudaMemcpy(h, tp.trace, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[tc3 trace BN=%d Kt=%d %dx%d stages=%d stationary=%d] producer: wait_empty %llu of %llu | mma0 (%llu pairs): wait_acc %llu wait_full %llu "
                "of %llu | mma1: wait_acc %llu wait_full %llu of %llu | epilogue: wait_acc_full %llu of %llu cycles\n", BN, tp.kchunks * KS, tp.H, tp.W,
                tp.stages, tp.stationary, h[0], h[1], h[5], h[2], h[3], h[4], h[6], h[7], h[8], h[10], h[11]);
    }
    return 0;
}

}  // namespace

bool vv_igemm_tc3_supported(const VvIGemm &p) {
    static int off = -1;
    if (off < 0) {
        const char *e = getenv("VECVAD_NO_TC3");
        off = (e && e[0] == '1') ? 1 : 0;
    }
    if (off || !vv_igemm_tc_supported(p)) return false;
    Tc3Params tp;
    int smem;
    memset(&tp, 0, sizeof(tp));
    return plan3(p, tp, smem);
}

int vv_launch_igemm_tc3(const VvIGemm &p, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    Tc3Params tp;
    int smem = 0;
    memset(&tp, 0, sizeof(tp));
    VV_REQUIRE(enc && plan3(p, tp, smem), "igemm_tc3: unsupported shape (Kt=%d N=%d H=%d W=%d taps=%d)", p.Kt, p.N, p.H, p.W, p.taps.n);
    tp.N = p.N; tp.O = p.O; tp.o_gs = p.o_gs; tp.ldo = p.ldo; tp.o_coff = p.o_coff; tp.o_d2s = p.o_d2s;
    tp.bias = p.bias; tp.bias_gs = p.bias_gs; tp.stats = p.stats; tp.stats_gs = p.stats_gs;
    tp.o_f16 = p.o_f16; tp.o_split = p.o_split; tp.O2 = (__half *)p.O2; tp.o2_gs = p.o2_gs; tp.ldo2 = p.ldo2; tp.rev = p.rev;
    {
        static int tr = -1;
        static unsigned long long *buf = nullptr;
        if (tr < 0) { const char *e = getenv("VECVAD_TC3_TRACE"); tr = e ? atoi(e) : 0; }
        if (tr && !buf) VV_CK(cudaMalloc(&buf, 16 * sizeof(unsigned long long)));
        tp.trace = tr ? buf : nullptr;
    }
    const int rows = tp.bh + tp.ndy - 1;
    const int bn_tile = tc3_bn_tile(p.N);
    const int esz = p.ab_f16 ? 2 : 4;
    const CUtensorMapDataType dt = p.ab_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : tmap_dtype();
    const CUtensorMapSwizzle sw = p.ab_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;      // 32 channels = 64 / 128 bytes
    alignas(64) CUtensorMap tmA, tmB;
    {
        // dimensions ordered (channel, x, image, y): the box lands in shared memory as [row][image][x][32 ch]
        const int sc = p.a_s2d ? 2 : 1;
        const cuuint64_t C = p.a_s2d ? p.Kt / 4 : p.Kt;
        cuuint64_t dims[4] = {C, (cuuint64_t)sc * p.W, (cuuint64_t)p.G * p.B, (cuuint64_t)sc * p.H};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * esz, (cuuint64_t)sc * p.H * sc * p.W * p.lda * esz, (cuuint64_t)sc * p.W * p.lda * esz};
        cuuint32_t box[4] = {KS, (cuuint32_t)(sc * tp.bw), (cuuint32_t)tp.bn, (cuuint32_t)(sc * rows)};
        cuuint32_t estr[4] = {1, (cuuint32_t)sc, 1, (cuuint32_t)sc};
        CUresult r = enc(&tmA, dt, 4, (void *)((const char *)p.A + (long long)p.a_coff * esz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_tc3: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Kt, (cuuint64_t)p.N, (cuuint64_t)p.taps.n * p.G};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kt * esz, (cuuint64_t)p.N * p.Kt * esz};
        cuuint32_t box[3] = {KS, (cuuint32_t)bn_tile, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmB, dt, 3, (void *)p.Wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_tc3: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    const int n_tiles = p.N / bn_tile;
    int gx = 148 / (n_tiles * p.G);
    if (gx < 1) gx = 1;
    if (gx > tp.pairs) gx = tp.pairs;
    dim3 grid(gx, n_tiles, p.G);
    if (p.ab_f16) {
        if (bn_tile == 128) return launch3<128, true>(tmA, tmB, tp, grid, smem, st);
        if (bn_tile == 64) return launch3<64, true>(tmA, tmB, tp, grid, smem, st);
        return launch3<32, true>(tmA, tmB, tp, grid, smem, st);
    }
    if (bn_tile == 128) return launch3<128, false>(tmA, tmB, tp, grid, smem, st);
    if (bn_tile == 64) return launch3<64, false>(tmA, tmB, tp, grid, smem, st);
    return launch3<32, false>(tmA, tmB, tp, grid, smem, st);
}
