// tcgen05 (kind::tf32) 3x3 conv tiles over a FLATTENED, zero-separated pixel sequence: ONE activation box serves all 9 taps.
//
//   out[m, n] = bias[n] + sum_t sum_k A[shift(m, t), k] * Wt[t][n][k]          (VvIGemm, common.h; 3x3 taps of either sign)
//
// The pair tiles (igemm_tc3.cu) load one box per distinct dx because a dx step inside a [rows][W] box wraps around the image
// edge: at 32x32 resolution that is 3 boxes of 6 rows per 4 output rows = 4.5 x the activation bytes through L2 -> shared
// memory.  Here the box is (W + 1) pixels wide: TMA zero-fills the out-of-range column x = W, so in shared memory image rows
// sit P = W + 1 pixel-rows apart with ONE zero pixel between them, which is both the right padding of row y and the left
// padding of row y + 1.  In that flattened sequence every tap (dy, dx) of every pixel is the constant offset dy * P + dx, so
// the A operand of a tap is the SAME box read through a UMMA descriptor that starts (dy * P + dx) * 128 bytes further in --
// 9 taps from one load.  A tile is 128 consecutive sequence positions of one image (positions on the separator column or past
// the image are computed and discarded: 89 % useful rows at 32x32); its box is the 7 image rows those positions and their
// halo touch: 2.0 x the activation bytes instead of 4.5 x.
//   Descriptor starts are 128-byte but not 1024-byte aligned.  Measured on the device (scratch/flat_probe.py): the UMMA applies
//   the 128B swizzle to the absolute shared-memory address, so such a start reads the TMA-written box correctly with the
//   descriptor's base-offset field left at 0 (setting it to the start row's phase gives wrong operands).
//
// Roles (352 threads): warp 0 TMA producer | warps 1-2 MMA issuers, alternating tiles, each with its own accumulator and its
// own ring of TMA stages (one warp issues an MMA every ~75 cycles, the tensor pipe retires an N = 32 MMA in 40) | warps 3-6 and
// 7-10: two epilogue sets, one per accumulator: tcgen05.ld, release the accumulator at once, then bias, BatchNorm statistics
// and NHWC stores through a padded shared-memory staging tile (column sums without shuffles, full 128-byte lines per store
// instruction).  At N = 32 the UMMA's operand reads take every shared-memory cycle (4 KB of A + 1 KB of B per 40-cycle MMA), so
// everything the epilogue does through shared memory is slow (~2300 cycles per 128 x 32 block, measured): two sets keep it off
// the critical path.
#include "tc_common.cuh"

#ifndef VECVAD_FLAT_DEFAULT
#define VECVAD_FLAT_DEFAULT 1
#endif

namespace {

struct FlatParams {
    int B, H, W, G;
    int P, L, tpi, m_tiles;         // row pitch W+1, positions per image H*P, tiles per image, tiles in total
    unsigned mP, mtpi;              // ceil(2^32 / P), ceil(2^32 / tpi): n / d == __umulhi(n, m) for n * d < 2^32
    int rows;                       // image rows per box
    int kchunks, spr;               // 32-channel slabs per tile; TMA stages per ring (two rings, one per MMA warp)
    int tap[3][3];                  // [dy+1][dx+1] -> tap index of the weight tensor
    int a_bytes, stage_bytes;
    unsigned long long *trace;      // VECVAD_FLAT_TRACE: cycle counters of CTA (0,0,0), see launch_flat
    int N;
    float *O;
    long long o_gs;
    int ldo, o_coff;
    const float *bias;
    long long bias_gs;
    double *stats;
    long long stats_gs;
    int o_split;                    // columns >= o_split go to O2 as fp16 (VvIGemm::o_split); 0 = none
    __half *O2;
    long long o2_gs;
    int ldo2;
    int o_f16;                      // O itself is fp16 (raw conv outputs of the fp16 mode)
    int rev;                        // walk the tiles from the last to the first
};

constexpr int FL_THREADS = 352;
constexpr int FL_SMEM_MAX = 227 * 1024;
constexpr int FL_GUARD = 128;       // zeroed bytes in front of stage 0 (a tile starting at x = 0 reads one pixel-row before its box)
constexpr int FL_STG_LD = 36;       // floats per staging row: 16-byte aligned, conflict-free for row writes and column reads
constexpr int FL_STG_BYTES = 32 * FL_STG_LD * 4;
constexpr int FL_MAX_RING = 4;

// F16: operands are fp16 in HBM and shared memory (64-byte pixel rows, SWIZZLE_64B, kind::f16 with K = 16 per MMA): the same
// bytes per MMA carry twice the K, so half the MMAs, half the TMA bytes.  Accumulation and outputs stay fp32.
template <int BN, bool F16>
__global__ void __launch_bounds__(FL_THREADS, 1) k_igemm_flat(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                              const FlatParams p) {
    constexpr int ROWB = F16 ? KS * 2 : KS * 4;              // bytes of one pixel row of a 32-channel slab
    constexpr int KSTEPS = F16 ? KS / 16 : KS / 8;           // MMAs per (tap, slab)
    constexpr int B_TAP = BN * ROWB;                         // one (tap, slab) weight tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *stage0 = smem_raw + FL_GUARD + vv_smem_pad(smem_raw + FL_GUARD, 1024);   // >= FL_GUARD bytes of ours in front
                                                             // stage (ring r, slot j) at stage0 + (r * spr + j) * stage_bytes
    uint8_t *b_stat = stage0 + 2 * p.spr * p.stage_bytes;
    uint8_t *stg_base = b_stat + 9 * p.kchunks * B_TAP;
    uint8_t *tail = stg_base + 8 * FL_STG_BYTES;
    uint64_t *full = (uint64_t *)tail;                       // [2][FL_MAX_RING]
    uint64_t *empty = full + 2 * FL_MAX_RING;                // [2][FL_MAX_RING]
    uint64_t *acc_full = empty + 2 * FL_MAX_RING;            // [2]
    uint64_t *acc_empty = acc_full + 2;                      // [2]
    uint64_t *bfull = acc_empty + 2;                         // [1]
    uint32_t *tmem_slot = (uint32_t *)(bfull + 1);
    float *s_bias = (float *)(tail + 256);
    float *s_sum = s_bias + BN, *s_sq = s_sum + BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    const bool tracer = p.trace && blockIdx.x + blockIdx.y + blockIdx.z == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * FL_MAX_RING; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
    // zero what TMA never writes but the MMAs may read: the guard in front of stage 0 and the slack behind every box
    for (int i = threadIdx.x; i < FL_GUARD / 16; i += FL_THREADS) reinterpret_cast<uint4 *>(stage0 - FL_GUARD)[i] = make_uint4(0, 0, 0, 0);
    {
        const int slack16 = (p.stage_bytes - p.a_bytes) / 16;
        for (int i = threadIdx.x; i < 2 * p.spr * slack16; i += FL_THREADS) {
            const int s = i / slack16, j = i - s * slack16;
            reinterpret_cast<uint4 *>(stage0 + s * p.stage_bytes + p.a_bytes)[j] = make_uint4(0, 0, 0, 0);
        }
    }
    vv_pdl_wait();                                           // set-up above overlaps the previous kernel's tail; global memory from here on
    for (int i = threadIdx.x; i < BN; i += FL_THREADS) {
        s_bias[i] = p.bias ? p.bias[g * p.bias_gs + n0 + i] : 0.f;
        s_sum[i] = 0.f; s_sq[i] = 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy zeros -> visible to the UMMA (async proxy) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: weights once, then one box per (tile, 32-channel slab) into the ring of the MMA warp
            // that owns the tile (tile parity)
            mbar_expect_tx(bfull, 9 * p.kchunks * B_TAP);
            for (int t = 0; t < 9; t++)
                for (int kc = 0; kc < p.kchunks; kc++)
                    tma_load_3d(b_stat + (t * p.kchunks + kc) * B_TAP, &tmB, bfull, kc * KS, n0, g * 9 + t);
            int slot0 = 0, slot1 = 0, round0 = 0, round1 = 0;
            long long t_wait = 0;
            const long long t_begin = clock64();
            int tcount = 0;
            for (int lt = blockIdx.x; lt < p.m_tiles; lt += gridDim.x, tcount++) {
                const int tile = p.rev ? p.m_tiles - 1 - lt : lt;
                const int r = tcount & 1;
                const int img = (p.tpi == 1 ? tile : (int)__umulhi((unsigned)tile, p.mtpi)), tt = tile - img * p.tpi;
                const int r_lo = (int)__umulhi((unsigned)(tt * BM), p.mP);                      // first image row with a position in this tile
                int slot = r ? slot1 : slot0, round = r ? round1 : round0;
                for (int kc = 0; kc < p.kchunks; kc++) {
                    const int s = r * FL_MAX_RING + slot;
                    if (round > 0) {
                        const long long t0 = clock64();
                        mbar_wait(&empty[s], (round - 1) & 1);
                        t_wait += clock64() - t0;
                    }
                    mbar_expect_tx(&full[s], p.a_bytes);
                    tma_load_4d(stage0 + (r * p.spr + slot) * p.stage_bytes, &tmA, &full[s], kc * KS, 0, r_lo - 1, g * p.B + img);
                    if (++slot == p.spr) { slot = 0; round++; }
                }
                if (r) { slot1 = slot; round1 = round; } else { slot0 = slot; round0 = round; }
            }
            if (tracer) { p.trace[0] = t_wait; p.trace[1] = clock64() - t_begin; }
        }
    } else if (warp <= 2) {
        // ---------------- MMA issuers: warp 1 takes this CTA's even tiles (accumulator 0, ring 0), warp 2 the odd ones.
        // The loop is warp-uniform and free of divisions per MMA; one elected lane issues.
        const int mw = warp - 1;
        const uint32_t idesc = F16 ? idesc_f16(BN) : idesc_tf32(BN);
        const uint32_t ring0 = smem_u32(stage0) + mw * p.spr * p.stage_bytes;
        const uint32_t b_lo0 = (smem_u32(b_stat) & 0x3FFFF) >> 4;
        // 8-row groups 8 * ROWB bytes apart, version 1, SWIZZLE_128B (layout 2) / SWIZZLE_64B (layout 4)
        const uint32_t desc_hi = (uint32_t)((8 * ROWB) >> 4) | (1u << 14) | ((F16 ? 4u : 2u) << 29);
        const uint32_t d_tmem = tmem + mw * BN;
        uint64_t *rfull = full + mw * FL_MAX_RING, *rempty = empty + mw * FL_MAX_RING;
        long long t_wacc = 0, t_wfull = 0, t_wb;
        const long long t_begin = clock64();
        mbar_wait(bfull, 0);
        t_wb = clock64() - t_begin;
        int use = 0, s = 0, ph = 0;
        for (int lt = blockIdx.x + mw * gridDim.x; lt < p.m_tiles; lt += 2 * gridDim.x, use++) {
            const int tile = p.rev ? p.m_tiles - 1 - lt : lt;
            const int img = (p.tpi == 1 ? tile : (int)__umulhi((unsigned)tile, p.mtpi)), tt = tile - img * p.tpi;
            const int r_lo = (int)__umulhi((unsigned)(tt * BM), p.mP);
            const int q0 = tt * BM - r_lo * p.P + p.P;                 // box row of the tile's first position (box row 0 = image row r_lo-1, x=0)
            if (use > 0) {
                const long long t0 = clock64();
                mbar_wait(&acc_empty[mw], (use - 1) & 1);
                t_wacc += clock64() - t0;
            }
            tc_fence_after();
            uint32_t acc = 0;
            for (int kc = 0; kc < p.kchunks; kc++) {
                {
                    const long long t0 = clock64();
                    mbar_wait(&rfull[s], ph);
                    t_wfull += clock64() - t0;
                }
                tc_fence_after();
                // descriptor low words (address >> 4); rows are ROWB / 16 units; row offset >= -1: the guard / the previous stage's slack
                const uint32_t a_lo0 = (((ring0 + s * p.stage_bytes) & 0x3FFFF) >> 4) + (uint32_t)((q0 - p.P - 1) * (ROWB / 16));
                const uint32_t b_lok = b_lo0 + (uint32_t)(kc * (B_TAP >> 4));
#pragma unroll
                for (int dyi = 0; dyi < 3; dyi++) {
#pragma unroll
                    for (int dxi = 0; dxi < 3; dxi++) {
                        const uint32_t a_lo = a_lo0 + (uint32_t)((dyi * p.P + dxi) * (ROWB / 16));
                        const uint32_t b_lo = b_lok + (uint32_t)(p.tap[dyi][dxi] * p.kchunks * (B_TAP >> 4));
#pragma unroll
                        for (int k = 0; k < KSTEPS; k++) {       // 32 bytes of K per MMA in either format
                            const uint64_t da = ((uint64_t)desc_hi << 32) | (a_lo + 2 * k | (1u << 16));
                            const uint64_t db = ((uint64_t)desc_hi << 32) | (b_lo + 2 * k | (1u << 16));
                            if (elect_one()) {
                                if (F16) tc_mma_f16(d_tmem, da, db, idesc, (acc | (uint32_t)(dyi | dxi | k)) ? 1u : 0u);
                                else tc_mma_tf32(d_tmem, da, db, idesc, (acc | (uint32_t)(dyi | dxi | k)) ? 1u : 0u);
                            }
                        }
                    }
                }
                acc = 1;
                if (elect_one()) tc_commit(&rempty[s]);
                __syncwarp();
                if (++s == p.spr) { s = 0; ph ^= 1; }
            }
            if (elect_one()) tc_commit(&acc_full[mw]);
            __syncwarp();
        }
        if (tracer && lane == 0) {
            unsigned long long *tr = p.trace + 2 + 5 * mw;
            tr[0] = t_wacc; tr[1] = t_wfull; tr[2] = clock64() - t_begin; tr[3] = use; tr[4] = t_wb;
        }
    } else {
        // ---------------- epilogue warps 3..6 (set 0: even tiles, accumulator 0) and 7..10 (set 1); warp w may touch TMEM lanes
        // 32*(w%4) .. +31 (= tile rows)
        const int q = warp & 3, es = (warp - 3) >> 2;
        const int row = q * 32 + lane;
        float *O = p.O + g * p.o_gs + p.o_coff + n0;
        float *stg = (float *)(stg_base + (warp - 3) * FL_STG_BYTES);
        const int HW = p.H * p.W;
        long long t_wfull = 0, t_ld = 0, t_stage = 0, t_stat = 0, t_store = 0;
        float c_sum[BN / 32], c_sq[BN / 32];                 // lane j: running sums of columns c0 + j over this warp's rows
#pragma unroll
        for (int c = 0; c < BN / 32; c++) { c_sum[c] = 0.f; c_sq[c] = 0.f; }
        const long long t_begin = clock64();
        int use = 0;
        for (int lt = blockIdx.x + es * gridDim.x; lt < p.m_tiles; lt += 2 * gridDim.x, use++) {
            const int tile = p.rev ? p.m_tiles - 1 - lt : lt;
            const int buf = es;
            const int img = (p.tpi == 1 ? tile : (int)__umulhi((unsigned)tile, p.mtpi)), tt = tile - img * p.tpi;
            const int pos = tt * BM + row;
            const int y = (int)__umulhi((unsigned)pos, p.mP), x = pos - y * p.P;
            const bool valid = x < p.W && y < p.H;                    // not the separator column, not past the image
            const int pix = img * HW + pos - y;                       // output pixel index: y * W + x = pos - y
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            {
                const long long t0 = clock64();
                mbar_wait(&acc_full[buf], use & 1);
                t_wfull += clock64() - t0;
            }
            tc_fence_after();
            long long t1 = clock64();
            float v[BN];
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) tc_ld32_nowait(tmem + ((uint32_t)(q * 32) << 16) + buf * BN + c0, v + c0);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);       // accumulator quarter is in registers: the MMA warp may reuse it
            { const long long t2 = clock64(); t_ld += t2 - t1; t1 = t2; }
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                // bias, then this warp's 32 x 32 block into the staging tile (rows of invalid positions as zeros)
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(s_bias + c0 + j);
                    float4 o;
                    o.x = valid ? v[c0 + j] + b4.x : 0.f;
                    o.y = valid ? v[c0 + j + 1] + b4.y : 0.f;
                    o.z = valid ? v[c0 + j + 2] + b4.z : 0.f;
                    o.w = valid ? v[c0 + j + 3] + b4.w : 0.f;
                    *reinterpret_cast<float4 *>(stg + lane * FL_STG_LD + j) = o;
                }
                __syncwarp();
                { const long long t2 = clock64(); t_stage += t2 - t1; t1 = t2; }
                if (p.stats) {          // lane j sums column c0 + j over the 32 rows
                    float s = 0.f, sq = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; r++) {
                        const float xv = stg[r * FL_STG_LD + lane];
                        s += xv;
                        sq = fmaf(xv, xv, sq);
                    }
                    c_sum[c0 / 32] += s;
                    c_sq[c0 / 32] += sq;
                }
                { const long long t2 = clock64(); t_stat += t2 - t1; t1 = t2; }
                if (p.o_f16 || (p.o_split && n0 + c0 >= p.o_split)) {       // fp16 destination: O itself, or the side output of this block
                    const bool side = p.o_split && n0 + c0 >= p.o_split;
                    __half *Oh = side ? p.O2 + g * p.o2_gs + (n0 + c0 - p.o_split) + 4 * (lane & 7)
                                      : reinterpret_cast<__half *>(p.O) + g * p.o_gs + p.o_coff + n0 + c0 + 4 * (lane & 7);
                    const int ldh = side ? p.ldo2 : p.ldo;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = 4 * i + (lane >> 3);
                        const float4 o = *reinterpret_cast<const float4 *>(stg + r * FL_STG_LD + 4 * (lane & 7));
                        const int rp = __shfl_sync(0xffffffffu, pix, r);
                        if ((vmask >> r) & 1) *reinterpret_cast<uint2 *>(Oh + (long long)rp * ldh) = pack_half4(o);
                    }
                } else {                                  // store instruction i: rows 4i .. 4i+3, eight lanes per 128-byte pixel row
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = 4 * i + (lane >> 3);
                        const float4 o = *reinterpret_cast<const float4 *>(stg + r * FL_STG_LD + 4 * (lane & 7));
                        const int rp = __shfl_sync(0xffffffffu, pix, r);
                        if ((vmask >> r) & 1) *reinterpret_cast<float4 *>(O + (long long)rp * p.ldo + c0 + 4 * (lane & 7)) = o;
                    }
                }
                __syncwarp();
                { const long long t2 = clock64(); t_store += t2 - t1; t1 = t2; }
            }
        }
        if (tracer && threadIdx.x == 96) {
            p.trace[12] = t_wfull; p.trace[13] = clock64() - t_begin; p.trace[14] = t_ld; p.trace[15] = t_stage; p.trace[16] = t_stat; p.trace[17] = t_store;
        }
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

inline int flat_bn_tile(int N) { return N % 64 == 0 ? 64 : 32; }

// VECVAD_FLAT=0 keeps the net engine on the per-dx-box pair tiles (igemm_tc3.cu)
int flat_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECVAD_FLAT");
        v = e ? atoi(e) : VECVAD_FLAT_DEFAULT;
    }
    return v;
}

template <int BN, bool F16>
int launch_flat(const CUtensorMap &tmA, const CUtensorMap &tmB, const FlatParams &fp, dim3 grid, int smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_igemm_flat<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SMEM_MAX));
        attr = true;
    }
    vv_launch(k_igemm_flat<BN, F16>, dim3(grid), dim3(FL_THREADS), smem, st, tmA, tmB, fp);
    VV_CKL();
    if (fp.trace) {      // debugging aid: synchronous
        unsigned long long h[18];
        VV_CK(cudaStreamSynchronize(st));
        VV_CK(cudaMemcpy(h, fp.trace, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[flat trace BN=%d Kt=%d] producer: wait_empty %llu of %llu | mma0 (%llu tiles): wait_weights %llu wait_acc %llu wait_full %llu of %llu | "
                "mma1 (%llu tiles): wait_acc %llu wait_full %llu of %llu | epilogue: wait_acc_full %llu of %llu cycles (ld %llu stage %llu stats %llu stores %llu)\n",
                BN, fp.kchunks * KS, h[0], h[1], h[5], h[6], h[2], h[3], h[4], h[10], h[7], h[8], h[9], h[12], h[13], h[14], h[15], h[16], h[17]);
    }
    return 0;
}

}  // namespace

// shapes the flattened-sequence tiles take: 3x3 taps, plain NHWC in/out, all nine weight tiles resident (<= 72 KB)
bool vv_igemm_flat_shape_ok(const VvIGemm &p) {
    FlatParams fp;
    if (!vv_igemm_tc_supported(p) || p.a_s2d || p.o_d2s || !analyse_3x3(p.taps, fp)) return false;
    if (p.o_f16 && (p.ldo % 8 || p.o_coff % 8)) return false;
    if (p.W + 1 > 256 || p.W < 8) return false;
    if (p.ab_f16 && (p.lda % 8 || p.a_coff % 8)) return false;           // 16-byte aligned pixel rows
    const int b_all = 9 * (p.Kt / KS) * flat_bn_tile(p.N) * KS * (p.ab_f16 ? 2 : 4);
    return b_all <= 72 * 1024;
}

// used by the net engine: only where the flattened tiles beat the per-dx boxes (full-width rows, i.e. W >= 32)
bool vv_igemm_flat_supported(const VvIGemm &p) { return flat_mode() != 0 && p.W >= 32 && vv_igemm_flat_shape_ok(p); }

int vv_launch_igemm_flat(const VvIGemm &p, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    FlatParams fp;
    memset(&fp, 0, sizeof(fp));
    VV_REQUIRE(enc && vv_igemm_flat_shape_ok(p) && analyse_3x3(p.taps, fp), "igemm_flat: unsupported shape (Kt=%d N=%d H=%d W=%d)", p.Kt, p.N,
               p.H, p.W);
    fp.B = p.B; fp.H = p.H; fp.W = p.W; fp.G = p.G;
    fp.P = p.W + 1; fp.L = p.H * fp.P; fp.tpi = (fp.L + BM - 1) / BM; fp.m_tiles = fp.tpi * p.B;
    fp.mP = (unsigned)((0x100000000ULL + fp.P - 1) / fp.P); fp.mtpi = (unsigned)((0x100000000ULL + fp.tpi - 1) / fp.tpi);
    fp.rows = (BM - 1) / fp.P + 4;                    // rows holding 128 consecutive positions (<= 127/P + 2) + one halo row either side
    const int esz = p.ab_f16 ? 2 : 4;                  // operand element size
    fp.kchunks = p.Kt / KS;
    fp.a_bytes = fp.rows * fp.P * KS * esz;
    fp.stage_bytes = (fp.a_bytes + KS * esz + 1023) / 1024 * 1024;    // >= one zero pixel-row of slack: the next stage's leading guard
    fp.N = p.N; fp.O = p.O; fp.o_gs = p.o_gs; fp.ldo = p.ldo; fp.o_coff = p.o_coff;
    fp.bias = p.bias; fp.bias_gs = p.bias_gs; fp.stats = p.stats; fp.stats_gs = p.stats_gs;
    fp.o_split = p.o_split; fp.O2 = (__half *)p.O2; fp.o2_gs = p.o2_gs; fp.ldo2 = p.ldo2; fp.o_f16 = p.o_f16; fp.rev = p.rev;
    const int bn_tile = flat_bn_tile(p.N);
    const int b_all = 9 * fp.kchunks * bn_tile * KS * esz;
    const int fixed = 1024 /*alignment*/ + FL_GUARD + 8 * FL_STG_BYTES + 256 /*barriers*/ + 3 * bn_tile * 4;
    int spr = (FL_SMEM_MAX - fixed - b_all) / (2 * fp.stage_bytes);
    {
        // tf32: the MMA warps are the limiter, deeper rings change nothing (measured); fp16: half the MMA time per box, so two
        // stages per ring no longer cover the TMA latency (52.9 -> 39.1 us at four, profiles/r01_flat16_probe.txt)
        static int cap = -1;
        if (cap < 0) { const char *e = getenv("VECVAD_FLAT_STAGES"); cap = e ? atoi(e) : 0; }
        const int c = cap > 0 ? cap : (p.ab_f16 ? 4 : 2);
        if (spr > c) spr = c;
    }
    if (spr > FL_MAX_RING) spr = FL_MAX_RING;
    VV_REQUIRE(spr >= 1, "igemm_flat: tile does not fit in shared memory");
    fp.spr = spr;
    {
        static int tr = -1;
        static unsigned long long *buf = nullptr;
        if (tr < 0) { const char *e = getenv("VECVAD_FLAT_TRACE"); tr = e ? atoi(e) : 0; }
        if (tr && !buf) VV_CK(cudaMalloc(&buf, 32 * sizeof(unsigned long long)));
        fp.trace = tr ? buf : nullptr;
    }
    const int smem = fixed + b_all + 2 * spr * fp.stage_bytes;

    const CUtensorMapDataType dt = p.ab_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : tmap_dtype();
    const CUtensorMapSwizzle sw = p.ab_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;      // 32 channels = 64 / 128 bytes
    const char *abase = (const char *)p.A + (long long)p.a_coff * esz;
    alignas(64) CUtensorMap tmA, tmB;
    {
        // (channel, x, y, image); the box is P = W + 1 wide and starts at x = 0: column W is out of range = the zero separator
        cuuint64_t dims[4] = {(cuuint64_t)p.Kt, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.G * p.B};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * esz, (cuuint64_t)p.W * p.lda * esz, (cuuint64_t)p.H * p.W * p.lda * esz};
        cuuint32_t box[4] = {KS, (cuuint32_t)fp.P, (cuuint32_t)fp.rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, dt, 4, (void *)abase, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_flat: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Kt, (cuuint64_t)p.N, (cuuint64_t)9 * p.G};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kt * esz, (cuuint64_t)p.N * p.Kt * esz};
        cuuint32_t box[3] = {KS, (cuuint32_t)bn_tile, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmB, dt, 3, (void *)p.Wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_flat: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    const int n_tiles = p.N / bn_tile;
    int gx = 148 / (n_tiles * p.G);
    if (gx < 1) gx = 1;
    if (gx > fp.m_tiles) gx = fp.m_tiles;
    dim3 grid(gx, n_tiles, p.G);
    if (p.ab_f16) return bn_tile == 64 ? launch_flat<64, true>(tmA, tmB, fp, grid, smem, st) : launch_flat<32, true>(tmA, tmB, fp, grid, smem, st);
    if (bn_tile == 64) return launch_flat<64, false>(tmA, tmB, fp, grid, smem, st);
    return launch_flat<32, false>(tmA, tmB, fp, grid, smem, st);
}
