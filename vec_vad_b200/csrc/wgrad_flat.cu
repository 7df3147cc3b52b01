// Weight gradient of a 3x3 conv on tensor cores over FLATTENED, zero-separated pixel sequences: ONE tcgen05.mma covers all
// nine taps of a K-step.
//
//   dW[t][n][k] += sum_m Gd[m, n] * A[shift(m, t), k]                     (VvWGrad, common.h; t = (dy, dx) in {-1,0,1}^2)
//
// k_wgrad_tc2 (wgrad_tc2.cu) issues one MMA per (dx, 8 pixels; 16 with fp16 operands): M = (dy, 32 input channels), N = output channels; at 32 output
// channels that is 48 MMAs of N = 32 per 128 pixels and the kernel runs at the rate one warp can issue them (~67 cycles each,
// measured; issuing from two warps is slower for these MN-major operands).  Both shifts can ride on descriptor offsets when
// activations AND output gradients sit in shared memory as sequences with one zero pixel between image rows (the TMA boxes
// are W + 1 wide, the out-of-range column arrives as zeros; igemm_flat.cu): with q = p + dx
//      dW[(dy,dx)][n][k] = sum_q Gd[q - dx][n] * A[q + dy * P][k],          P = W + 1
//   * M operand: four 32-channel blocks of the activation box, P pixel-rows apart (leading-dimension byte offset = P * row
//     bytes): dy = -1, 0, +1 (+ a fourth block nobody reads);
//   * N operand: three 32-channel blocks of the gradient box ONE pixel-row apart (leading-dimension byte offset = one row):
//     dx = +1, 0, -1;
//   so per K-step ONE MMA (M = 128, N = 96) instead of three of N = 32, and one activation box per tile instead of three.
//   Pixel rows are 128 bytes (tf32, K = 8 positions per MMA) or 64 bytes (template F16: fp16 operands, K = 16 positions).  Zero separators make the wrapped-around neighbours of edge pixels vanish; positions past the image have zero gradient.
// A CTA owns one 32-input-channel slab, 32 output channels and a share of the 128-position tiles; one TMEM accumulator of 96
// columns, flushed to dW with fp32 reductions at the end.
// warp 0: TMA producer | warp 1: MMA issuer (+ TMEM alloc) | warps 0-2: epilogue (warp = dy).
#include "tc_common.cuh"

#ifndef VECVAD_WGRAD_FLAT_DEFAULT
#define VECVAD_WGRAD_FLAT_DEFAULT 1
#endif

namespace {

struct WgfParams {
    int B, H, W, G;
    int P, tpi, m_tiles, tiles_per_split;
    unsigned mP, mtpi;              // ceil(2^32 / P), ceil(2^32 / tpi)
    int kchunks, n_tiles;
    int tap[3][3];                  // [dy+1][dx+1] -> tap index of dW
    int N, Kt;
    int a_rows, g_rows;             // image rows per activation / gradient box
    int a_bytes, g_bytes, stage_bytes, stages;
    float *dW;
    long long dw_gs;
    unsigned long long *trace;
};

constexpr int WGF_SMEM_MAX = 227 * 1024;
constexpr int WGF_GUARD = 128;      // zeroed bytes in front of a gradient box: a tile starting at x = 0 reads one pixel-row before it
constexpr int WGF_NT = 32;          // output channels per CTA: N of the MMA = 3 * 32

// MN-major operand descriptor: tf32 = SWIZZLE_128B with 32-byte atoms (layout type 1), fp16 = SWIZZLE_64B (layout type 4);
// 4-row (tf32) / 8-row (fp16) groups 512 bytes apart in both (see wgrad_tc2.cu)
template <bool F16>
__device__ __forceinline__ uint64_t desc_mn_flat(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(F16 ? 4 : 1) << 61);
}
template <bool F16>
__device__ __forceinline__ uint32_t idesc_mnmn_flat(int n) {
    return (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <bool F16>
__global__ void __launch_bounds__(128, 1) k_wgrad_flat(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                                                       const WgfParams p) {
    constexpr int ROWB = F16 ? KS * 2 : KS * 4;             // bytes of one pixel row of a 32-channel slab
    constexpr int KPOS = F16 ? 16 : 8;                      // positions (K) per MMA: 1024 bytes of either operand
    extern __shared__ uint8_t smem_raw[];
    // stage s: [activation box | slack][gradient box | slack]; the slack behind a box (>= WGF_GUARD bytes) is zeroed once and never
    // written by TMA: the slack of the activation box is the leading guard of the gradient box behind it
    uint8_t *smem = smem_raw + vv_smem_pad(smem_raw, 1024);
    const int a_span = p.stage_bytes - ((p.g_bytes + WGF_GUARD + 1023) / 1024 * 1024);      // activation box + slack
    uint8_t *tail = smem + p.stages * p.stage_bytes;
    uint64_t *full = (uint64_t *)tail, *empty = full + 8, *accum = empty + 8;
    uint32_t *tmem_slot = (uint32_t *)(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int kc = blockIdx.x % p.kchunks, nt = blockIdx.x / p.kchunks;
    const int n0 = nt * WGF_NT;
    const int pt_begin = blockIdx.y * p.tiles_per_split;
    const int pt_end = min(p.m_tiles, pt_begin + p.tiles_per_split);
    const int ntiles = pt_end - pt_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the slack behind every box (guards); TMA never writes there
    for (int s = 0; s < p.stages; s++) {
        uint8_t *st = smem + s * p.stage_bytes;
        for (int i = p.a_bytes + threadIdx.x * 16; i < a_span; i += 128 * 16) *reinterpret_cast<uint4 *>(st + i) = make_uint4(0, 0, 0, 0);
        for (int i = a_span + p.g_bytes + threadIdx.x * 16; i < p.stage_bytes; i += 128 * 16) *reinterpret_cast<uint4 *>(st + i) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    vv_pdl_wait();                                           // set-up above overlaps the previous kernel's tail; global memory from here on

    if (ntiles > 0) {
        if (warp == 0 && lane == 0) {
            // ---------------- TMA producer: per tile one activation box (halo rows included) and one gradient box
            int s = 0, round = 0;
            long long t_wait = 0;
            const long long t_begin = clock64();
            for (int ti = 0; ti < ntiles; ti++) {
                const int tile = pt_begin + ti;
                const int img = p.tpi == 1 ? tile : (int)__umulhi((unsigned)tile, p.mtpi), tt = tile - img * p.tpi;
                const int r_lo = (int)__umulhi((unsigned)(tt * BM), p.mP);       // first image row with a position in this tile
                if (round > 0) { const long long t0 = clock64(); mbar_wait(&empty[s], (round - 1) & 1); t_wait += clock64() - t0; }
                uint8_t *st = smem + s * p.stage_bytes;
                mbar_expect_tx(&full[s], p.a_bytes + p.g_bytes);
                tma_load_5d(st, &tmA, &full[s], kc * KS, 0, r_lo - 1, img, g);
                tma_load_5d(st + a_span, &tmG, &full[s], n0, 0, r_lo, img, g);
                if (++s == p.stages) { s = 0; round++; }
            }
            if (p.trace && blockIdx.x + blockIdx.y + blockIdx.z == 0) { p.trace[0] = t_wait; p.trace[1] = clock64() - t_begin; }
        } else if (warp == 1) {
            // ---------------- MMA issuer (warp-uniform loop, one elected lane issues)
            const uint32_t idesc = idesc_mnmn_flat<F16>(3 * WGF_NT);
            const uint32_t base = smem_u32(smem);
            int s = 0, ph = 0;
            long long t_wait = 0;
            const long long t_begin = clock64();
            for (int ti = 0; ti < ntiles; ti++) {
                const int tile = pt_begin + ti;
                const int img = p.tpi == 1 ? tile : (int)__umulhi((unsigned)tile, p.mtpi), tt = tile - img * p.tpi;
                const int r_lo = (int)__umulhi((unsigned)(tt * BM), p.mP);
                const int g0 = tt * BM - r_lo * p.P;                 // gradient-box row of the tile's first position (box row 0 = image row r_lo, x = 0)
                { const long long t0 = clock64(); mbar_wait(&full[s], ph); t_wait += clock64() - t0; }
                tc_fence_after();
                // activation box row 0 = image row r_lo - 1: block dy starts at row g0 + (dy + 1) * P, i.e. block 0 (dy = -1) at g0
                const uint64_t da = desc_mn_flat<F16>(base + s * p.stage_bytes + g0 * ROWB, p.P * ROWB);
                // gradient block j = Gd[q + j - 1]  <->  dx = 1 - j; the first one starts one pixel-row before the tile
                const uint64_t dg = desc_mn_flat<F16>(base + s * p.stage_bytes + a_span + (g0 - 1) * ROWB, ROWB);
#pragma unroll
                for (int k = 0; k < BM / KPOS; k++)   // KPOS positions = 1024 bytes per MMA in both operands
                    if (elect_one()) {
                        if (F16) tc_mma_f16(tmem, da + 64 * k, dg + 64 * k, idesc, (ti | k) ? 1u : 0u);
                        else tc_mma_tf32(tmem, da + 64 * k, dg + 64 * k, idesc, (ti | k) ? 1u : 0u);
                    }
                if (elect_one()) tc_commit(&empty[s]);
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) tc_commit(accum);
            __syncwarp();
            if (p.trace && lane == 0 && blockIdx.x + blockIdx.y + blockIdx.z == 0) { p.trace[2] = t_wait; p.trace[3] = clock64() - t_begin; p.trace[4] = ntiles; }
        }
        __syncwarp();
        // ---------------- epilogue: warp = dy + 1, lane = input channel within the slab (contiguous in dW), column = (dx block, output channel)
        mbar_wait(accum, 0);
        tc_fence_after();
        if (warp < 3) {
#pragma unroll 1
            for (int j = 0; j < 3; j++) {
                const int t = p.tap[warp][2 - j];                    // block j holds dx = 1 - j
                float *dst = p.dW + g * p.dw_gs + ((long long)t * p.N + n0) * p.Kt + kc * KS + lane;
                float v[32];
                tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + j * WGF_NT, v);
#pragma unroll
                for (int c = 0; c < 32; c++) atomicAdd(dst + (long long)c * p.Kt, v[c]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

bool analyse_3x3_wg(const VvTaps &t, WgfParams &wp) {
    if (t.n != 9) return false;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) wp.tap[i][j] = -1;
    for (int k = 0; k < 9; k++) {
        if (t.dy[k] < -1 || t.dy[k] > 1 || t.dx[k] < -1 || t.dx[k] > 1) return false;
        wp.tap[t.dy[k] + 1][t.dx[k] + 1] = k;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (wp.tap[i][j] < 0) return false;
    return true;
}

int wgf_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECVAD_WGRAD_FLAT");
        v = e ? atoi(e) : VECVAD_WGRAD_FLAT_DEFAULT;
    }
    return v;
}

}  // namespace

bool vv_wgrad_flat_shape_ok(const VvWGrad &p) {
    WgfParams wp;
    if (!vv_wgrad_tc_supported(p) || p.g_s2d || !analyse_3x3_wg(p.taps, wp)) return false;
    if (p.W + 1 > 256 || p.W < 8 || p.N % WGF_NT) return false;
    return true;
}

// used by the net engine: where the single-MMA-per-K-step tiles win (full-width rows, 32 output channels)
bool vv_wgrad_flat_supported(const VvWGrad &p) { return wgf_mode() != 0 && p.W >= 32 && p.N == 32 && vv_wgrad_flat_shape_ok(p); }

int vv_launch_wgrad_flat(const VvWGrad &p, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    WgfParams wp;
    memset(&wp, 0, sizeof(wp));
    VV_REQUIRE(enc && vv_wgrad_flat_shape_ok(p) && analyse_3x3_wg(p.taps, wp), "wgrad_flat: unsupported shape (Kt=%d N=%d H=%d W=%d)", p.Kt, p.N, p.H,
               p.W);
    wp.B = p.B; wp.H = p.H; wp.W = p.W; wp.G = p.G;
    wp.P = p.W + 1;
    const int L = p.H * wp.P;
    wp.tpi = (L + BM - 1) / BM; wp.m_tiles = wp.tpi * p.B;
    wp.mP = (unsigned)((0x100000000ULL + wp.P - 1) / wp.P); wp.mtpi = (unsigned)((0x100000000ULL + wp.tpi - 1) / wp.tpi);
    wp.kchunks = p.Kt / KS; wp.n_tiles = p.N / WGF_NT;
    wp.N = p.N; wp.Kt = p.Kt; wp.dW = p.dW; wp.dw_gs = p.dw_gs;
    wp.g_rows = (BM - 1) / wp.P + 2;                 // rows holding 128 consecutive positions
    wp.a_rows = wp.g_rows + 2;                       // + one halo row either side
    const int esz = p.ab_f16 ? 2 : 4;
    wp.a_bytes = wp.a_rows * wp.P * KS * esz;
    wp.g_bytes = wp.g_rows * wp.P * KS * esz;
    const int a_span = (wp.a_bytes + WGF_GUARD + 1023) / 1024 * 1024, g_span = (wp.g_bytes + WGF_GUARD + 1023) / 1024 * 1024;
    wp.stage_bytes = a_span + g_span;
    // the fourth M block (dy = +2, never read back) starts 3 * P rows into the tile: keep its reads inside our allocation
    const int overread = (3 * wp.P + BM + wp.P) * KS * esz;
    const int fixed = 1024 + 256 + (overread > wp.stage_bytes ? overread - wp.stage_bytes : 0);
    int stages = (WGF_SMEM_MAX - fixed) / wp.stage_bytes;
    {
        static int cap = -1;
        if (cap < 0) { const char *e = getenv("VECVAD_WGF_STAGES"); cap = e ? atoi(e) : 3; }
        if (cap >= 2 && stages > cap) stages = cap;
    }
    if (stages > 8) stages = 8;
    VV_REQUIRE(stages >= 2, "wgrad_flat: tile does not fit in shared memory");
    wp.stages = stages;
    {
        static int tr = -1;
        static unsigned long long *buf = nullptr;
        if (tr < 0) { const char *e = getenv("VECVAD_WGF_TRACE"); tr = e ? atoi(e) : 0; }
        if (tr && !buf) VV_CK(cudaMalloc(&buf, 16 * sizeof(unsigned long long)));
        wp.trace = tr ? buf : nullptr;
    }
    const int smem = fixed + stages * wp.stage_bytes;

    const CUtensorMapDataType dt = p.ab_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : tmap_dtype();
    const CUtensorMapSwizzle sw = p.ab_f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    const unsigned long long e = esz;
    alignas(64) CUtensorMap tmA, tmG;
    {
        // (channel, x, y, image, group); the boxes are P = W + 1 wide from x = 0: column W is out of range = the zero separator
        cuuint64_t dims[5] = {(cuuint64_t)p.Kt, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B, (cuuint64_t)p.G};
        cuuint64_t strides[4] = {(cuuint64_t)p.lda * e, (cuuint64_t)p.W * p.lda * e, (cuuint64_t)p.H * p.W * p.lda * e,
                                 (cuuint64_t)(p.G > 1 ? p.a_gs : (long long)p.B * p.H * p.W * p.lda) * e};
        cuuint32_t box[5] = {KS, (cuuint32_t)wp.P, (cuuint32_t)wp.a_rows, 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmA, dt, 5, (void *)((const char *)p.A + (long long)p.a_coff * esz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "wgrad_flat: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[5] = {(cuuint64_t)p.N, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B, (cuuint64_t)p.G};
        cuuint64_t strides[4] = {(cuuint64_t)p.ldg * e, (cuuint64_t)p.W * p.ldg * e, (cuuint64_t)p.H * p.W * p.ldg * e,
                                 (cuuint64_t)(p.G > 1 ? p.g_gs : (long long)p.B * p.H * p.W * p.ldg) * e};
        cuuint32_t box[5] = {KS, (cuuint32_t)wp.P, (cuuint32_t)wp.g_rows, 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmG, dt, 5, (void *)((const char *)p.Gd + (long long)p.g_coff * esz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "wgrad_flat: cuTensorMapEncodeTiled(Gd) failed with %d", (int)r);
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
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_wgrad_flat<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WGF_SMEM_MAX));
        VV_CK(cudaFuncSetAttribute(k_wgrad_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WGF_SMEM_MAX));
        attr = true;
    }
    if (p.ab_f16) vv_launch(k_wgrad_flat<true>, dim3(grid), dim3(128), smem, st, tmA, tmG, wp);
    else vv_launch(k_wgrad_flat<false>, dim3(grid), dim3(128), smem, st, tmA, tmG, wp);
    VV_CKL();
    if (wp.trace) {      // debugging aid: synchronous
        unsigned long long h[5];
        VV_CK(cudaStreamSynchronize(st));
        VV_CK(cudaMemcpy(h, wp.trace, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[wgf trace Kt=%d %dx%d stages=%d grid=(%d,%d,%d) tiles/CTA=%llu] producer: wait_empty %llu of %llu | mma: wait_full %llu of %llu cycles\n",
                wp.Kt, wp.H, wp.W, wp.stages, grid.x, grid.y, grid.z, h[4], h[0], h[1], h[2], h[3]);
    }
    return 0;
}
