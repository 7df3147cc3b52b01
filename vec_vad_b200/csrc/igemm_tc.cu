// tcgen05 (kind::tf32) implicit-GEMM tiles for the conv / transposed-conv / input-gradient contractions of the
// completion UNets (reference: the cuDNN calls behind model/unet.py:9-16,54,66 and their autograd backward).
//
//   out[m, n] = bias[n] + sum_t sum_k A[shift(m, t), k] * Wt[t][n][k]          (VvIGemm, common.h)
//
// One CTA = 128 output pixels (a TMA box of bw x bh pixels of bn images) x BN output channels.
//   warp 0 / lane 0 : TMA producer.  Per (tap, 32-channel slab): one 4-D box load of the NHWC activations shifted by
//                     the tap (out-of-image pixels are zero-filled by TMA = the conv's zero padding) + one 3-D box
//                     load of the weights, both SWIZZLE_128B, into a ring of smem stages.
//   warp 1 / lane 0 : issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) x4 per stage, accumulating over
//                     all taps and slabs in TMEM; tcgen05.commit frees the stage / publishes the accumulator.
//   warps 0-3       : epilogue. tcgen05.ld 32 lanes x 32 columns per warp, + bias, store NHWC (optionally pixel-
//                     shuffled for the transposed conv), BatchNorm batch statistics by a shuffle butterfly.
// fp32 activations are fed to the tensor core as-is (kind::tf32 reads the top 19 bits); accumulation, bias,
// statistics and everything downstream stay fp32.
#include "tc_common.cuh"

namespace {

constexpr int A_STAGE = BM * KS * 4;

struct TcParams {
    int B, H, W, G;
    int bw, bh, bn;                 // pixel box of one CTA tile: bw * bh * bn == 128
    int tiles_x, tiles_y, tiles_n;  // tiles per group = tiles_n * tiles_y * tiles_x
    int ntaps, kchunks, cq;         // cq: channels per space-to-depth phase (a_s2d), else 0
    int dy[9], dx[9];
    int N;
    float *O;
    long long o_gs;
    int ldo, o_coff, o_d2s;
    const float *bias;
    long long bias_gs;
    double *stats;
    long long stats_gs;
};

template <int BN, int STAGES>
struct Smem {
    static constexpr int B_STAGE = BN * KS * 4;
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int BYTES = STAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/ + 3 * BN * 4;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(128) k_igemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                  const TcParams p) {
    using SM = Smem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t *full = (uint64_t *)(smem + STAGES * SM::STAGE);
    uint64_t *empty = full + STAGES;
    uint64_t *accum = empty + STAGES;
    uint32_t *tmem_slot = (uint32_t *)(accum + 1);
    float *s_bias = (float *)(smem + STAGES * SM::STAGE + 256);
    float *s_sum = s_bias + BN, *s_sq = s_sum + BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    int tile = blockIdx.x;
    const int tx = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty = tile % p.tiles_y; tile /= p.tiles_y;
    const int img0 = tile * p.bn, y0 = ty * p.bh, x0 = tx * p.bw;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {   // TMEM allocation: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < BN; i += 128) {
        s_bias[i] = p.bias ? p.bias[g * p.bias_gs + (p.o_d2s ? (n0 + i) % (p.N >> 2) : (n0 + i))] : 0.f;
        s_sum[i] = 0.f; s_sq[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int nsteps = p.ntaps * p.kchunks;

    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        for (int it = 0; it < nsteps; it++) {
            const int s = it % STAGES, round = it / STAGES;
            if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
            const int t = it / p.kchunks, kc = it - t * p.kchunks;
            uint8_t *sa = smem + s * SM::STAGE, *sb = sa + A_STAGE;
            mbar_expect_tx(&full[s], SM::STAGE);
            int c = kc * KS, xx = x0 + p.dx[t], yy = y0 + p.dy[t];
            if (p.cq) {                     // space-to-depth source: channel slab -> (phase, channel), stride-2 pixel walk
                const int ph = c / p.cq;
                c -= ph * p.cq;
                xx = 2 * xx + (ph & 1);
                yy = 2 * yy + (ph >> 1);
            }
            tma_load_4d(sa, &tmA, &full[s], c, xx, yy, g * p.B + img0);
            tma_load_3d(sb, &tmB, &full[s], kc * KS, n0, g * p.ntaps + t);
        }
    } else if (warp == 1 && lane == 0) {
        // ---------------- MMA issuer
        const uint32_t idesc = idesc_tf32(BN);
        for (int it = 0; it < nsteps; it++) {
            const int s = it % STAGES, round = it / STAGES;
            mbar_wait(&full[s], round & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * SM::STAGE), sb = sa + A_STAGE;
            const uint64_t da = smem_desc_k_sw128(sa), db = smem_desc_k_sw128(sb);
#pragma unroll
            for (int k = 0; k < KS / 8; k++)      // 8 tf32 = 32 bytes per MMA along K: +2 in the (>>4) start-address field
                tc_mma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, (it | k) ? 1u : 0u);
            tc_commit(&empty[s]);                 // stage reusable once these MMAs have read it
        }
        tc_commit(accum);                          // accumulator complete
    }
    __syncwarp();

    // ---------------- epilogue (all four warps; warp w owns TMEM lanes 32w .. 32w+31 = tile rows)
    mbar_wait(accum, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    int r = row;
    const int xx = r % p.bw; r /= p.bw;
    const int yy = r % p.bh; r /= p.bh;
    const int b = img0 + r, y = y0 + yy, x = x0 + xx;
    const bool valid = b < p.B && y < p.H && x < p.W;
    float *O = p.O + g * p.o_gs;
    const int Co = p.o_d2s ? (p.N >> 2) : p.N;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] += s_bias[c0 + j];
        if (valid) {
            float *dst;
            const int ncol = n0 + c0;
            if (!p.o_d2s) {
                dst = O + ((long long)(b * p.H + y) * p.W + x) * p.ldo + p.o_coff + ncol;
            } else {             // N = 4*Co: column block (phase, co) -> pixel (2y+py, 2x+px); a 32-column block never straddles phases
                const int ph = ncol / Co, co = ncol - ph * Co;
                dst = O + ((long long)(b * 2 * p.H + 2 * y + (ph >> 1)) * (2 * p.W) + 2 * x + (ph & 1)) * p.ldo + p.o_coff + co;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (p.stats) {
            // column sums over this warp's 32 rows: butterfly transpose-reduce, 31 shuffles per quantity; lane j ends with column c0+j
            float s[32], q[32];
#pragma unroll
            for (int j = 0; j < 32; j++) { s[j] = valid ? v[j] : 0.f; q[j] = s[j] * s[j]; }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
                const bool up = lane & w;
#pragma unroll
                for (int j = 0; j < w; j++) {
                    float keep_s = up ? s[j + w] : s[j], send_s = up ? s[j] : s[j + w];
                    float keep_q = up ? q[j + w] : q[j], send_q = up ? q[j] : q[j + w];
                    s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
                    q[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
                }
            }
            atomicAdd(&s_sum[c0 + lane], s[0]);
            atomicAdd(&s_sq[c0 + lane], q[0]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.stats) {
        double *st = p.stats + g * p.stats_gs;
        for (int i = threadIdx.x; i < BN; i += 128) {
            if (n0 + i < p.N) {
                atomicAdd(&st[n0 + i], (double)s_sum[i]);
                atomicAdd(&st[p.N + n0 + i], (double)s_sq[i]);
            }
        }
    }
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
bool tile_geometry(int H, int W, int &bw, int &bh, int &bn) { return tile_geometry_n(H, W, BM, bw, bh, bn); }

int pick_bn(int N) { return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32); }

template <int BN, int STAGES>
int launch(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &tp, dim3 grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_igemm_tc<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN, STAGES>::BYTES));
        attr = true;
    }
    k_igemm_tc<BN, STAGES><<<grid, 128, Smem<BN, STAGES>::BYTES, st>>>(tmA, tmB, tp);
    VV_CKL();
    return 0;
}

}  // namespace

bool vv_igemm_tc_supported(const VvIGemm &p) {
    int bw, bh, bn;
    if (!encode_fn()) return false;
    if (p.Kt % KS || p.N % 32 || p.N < 32) return false;
    if (p.lda % 4 || p.a_coff % 4 || p.ldo % 4 || p.o_coff % 4) return false;
    if (((uintptr_t)p.A) % 16 || ((uintptr_t)p.Wt) % 16 || ((uintptr_t)p.O) % 16) return false;
    if (p.a_s2d && ((p.Kt / 4) % KS)) return false;
    if (p.o_d2s && ((p.N / 4) % 32)) return false;
    if (p.G > 1 && (p.a_gs != (long long)p.B * p.H * p.W * (p.a_s2d ? 4 : 1) * p.lda)) return false;   // groups must be contiguous images
    if (p.G > 1 && p.w_gs != (long long)p.taps.n * p.N * p.Kt) return false;
    if (!tile_geometry(p.H, p.W, bw, bh, bn)) return false;
    return true;
}

int vv_launch_igemm_tc(const VvIGemm &p, cudaStream_t st) {
    VV_REQUIRE(vv_igemm_tc_supported(p), "igemm_tc: unsupported shape (Kt=%d N=%d H=%d W=%d)", p.Kt, p.N, p.H, p.W);
    EncodeTiledFn enc = encode_fn();
    TcParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.B = p.B; tp.H = p.H; tp.W = p.W; tp.G = p.G;
    tile_geometry(p.H, p.W, tp.bw, tp.bh, tp.bn);
    tp.tiles_x = p.W / tp.bw; tp.tiles_y = p.H / tp.bh; tp.tiles_n = (p.B + tp.bn - 1) / tp.bn;
    tp.ntaps = p.taps.n; tp.kchunks = p.Kt / KS; tp.cq = p.a_s2d ? p.Kt / 4 : 0;
    for (int t = 0; t < 9; t++) { tp.dy[t] = p.taps.dy[t]; tp.dx[t] = p.taps.dx[t]; }
    tp.N = p.N; tp.O = p.O; tp.o_gs = p.o_gs; tp.ldo = p.ldo; tp.o_coff = p.o_coff; tp.o_d2s = p.o_d2s;
    tp.bias = p.bias; tp.bias_gs = p.bias_gs; tp.stats = p.stats; tp.stats_gs = p.stats_gs;

    const CUtensorMapDataType dt = tmap_dtype();
    alignas(64) CUtensorMap tmA, tmB;
    {
        const int sc = p.a_s2d ? 2 : 1;
        const cuuint64_t C = p.a_s2d ? p.Kt / 4 : p.Kt;
        cuuint64_t dims[4] = {C, (cuuint64_t)sc * p.W, (cuuint64_t)sc * p.H, (cuuint64_t)p.G * p.B};
        cuuint64_t strides[3] = {(cuuint64_t)p.lda * 4, (cuuint64_t)sc * p.W * p.lda * 4, (cuuint64_t)sc * p.H * sc * p.W * p.lda * 4};
        cuuint32_t box[4] = {KS, (cuuint32_t)(sc * tp.bw), (cuuint32_t)(sc * tp.bh), (cuuint32_t)tp.bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)sc, (cuuint32_t)sc, 1};
        CUresult r = enc(&tmA, dt, 4, (void *)(p.A + p.a_coff), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    const int bn_tile = pick_bn(p.N);
    {
        cuuint64_t dims[3] = {(cuuint64_t)p.Kt, (cuuint64_t)p.N, (cuuint64_t)p.taps.n * p.G};
        cuuint64_t strides[2] = {(cuuint64_t)p.Kt * 4, (cuuint64_t)p.N * p.Kt * 4};
        cuuint32_t box[3] = {KS, (cuuint32_t)bn_tile, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tmB, dt, 3, (void *)p.Wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "igemm_tc: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    dim3 grid(tp.tiles_x * tp.tiles_y * tp.tiles_n, p.N / bn_tile, p.G);
    if (bn_tile == 128) return launch<128, 3>(tmA, tmB, tp, grid, st);
    if (bn_tile == 64) return launch<64, 4>(tmA, tmB, tp, grid, st);
    return launch<32, 4>(tmA, tmB, tp, grid, st);
}

// ================================================================================================
// Weight gradient on tensor cores:   dW[t][n][k] += sum_m Gd[m, n] * A[shift(m, t), k]          (VvWGrad, common.h)
//
// The contraction runs over PIXELS, so both operands are fed MN-major straight from their NHWC home: a TMA box of
// 64 pixels x 32 channels is one "slab" (64 rows of 128 bytes, SWIZZLE_128B) and a tcgen05.mma (K = 8 pixels) reads
// eight rows of every slab (tf32 MN-major operands use the 128-byte swizzle with 32-byte atoms).  One CTA owns a 128 x NT block of the gradient:
//     rows    = 4 slabs of (tap, 32 input channels)   -- each slab is loaded with its own tap shift (zero-filled halo)
//     columns = NT/32 slabs of the output-channel gradient Gd
// accumulated in TMEM over its share of the pixel tiles (split over blockIdx.y), then added to dW with fp32 reductions.
// ================================================================================================
namespace {

constexpr int WPB = 64;                 // pixels per slab / pipeline stage
constexpr int W_SLAB = WPB * KS * 4;    // 8 KiB

// MN-major tf32 operands exist in one shared-memory layout only: 128-byte swizzle with 32-byte atomicity
// (UMMA layout type 1 = SWIZZLE_128B_BASE32B; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  32 channels (128 B) are contiguous,
// the swizzle pattern repeats every 4 pixel rows (512 B = SBO), 32-channel slabs are `lbo` bytes apart.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32_mn(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct WgParams {
    int B, H, W, G;
    int bw, bh, bn;                 // pixel box: bw * bh * bn == 64
    int tiles_x, tiles_y, tiles_n;  // pixel tiles per group
    int tiles_per_split;
    int ntaps, kchunks, nslabs;     // nslabs = ntaps * kchunks (rows of dW in units of 32)
    int dy[9], dx[9];
    int N, Kt, cq;                  // cq: channels per space-to-depth phase of Gd (g_s2d), else 0
    float *dW;
    long long dw_gs;
};

template <int NT, int STAGES>
struct WgSmem {
    static constexpr int NS = NT / 32;
    static constexpr int STAGE = (4 + NS) * W_SLAB;
    static constexpr int BYTES = STAGES * STAGE + 1024 + 256;
};

template <int NT, int STAGES>
__global__ void __launch_bounds__(128) k_wgrad_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                                                  const WgParams p) {
    using SM = WgSmem<NT, STAGES>;
    constexpr int NS = SM::NS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + STAGES * SM::STAGE);
    uint64_t *empty = full + STAGES;
    uint64_t *accum = empty + STAGES;
    uint32_t *tmem_slot = (uint32_t *)(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z;
    const int mgroups = (p.nslabs + 3) / 4;
    const int mg = blockIdx.x % mgroups, nt = blockIdx.x / mgroups;
    const int slab0 = mg * 4;
    const int mvalid = min(4, p.nslabs - slab0);
    const int n0 = nt * NT;
    const int tiles_total = p.tiles_x * p.tiles_y * p.tiles_n;
    const int pt_begin = blockIdx.y * p.tiles_per_split;
    const int pt_end = min(tiles_total, pt_begin + p.tiles_per_split);
    const int nsteps = pt_end - pt_begin;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (nsteps > 0) {
        if (warp == 0 && lane == 0) {
            // ---------------- TMA producer
            for (int it = 0; it < nsteps; it++) {
                const int s = it % STAGES, round = it / STAGES;
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                int tile = pt_begin + it;
                const int tx = tile % p.tiles_x; tile /= p.tiles_x;
                const int ty = tile % p.tiles_y; tile /= p.tiles_y;
                const int img0 = tile * p.bn, y0 = ty * p.bh, x0 = tx * p.bw;
                uint8_t *sa = smem + s * SM::STAGE, *sg = sa + 4 * W_SLAB;
                mbar_expect_tx(&full[s], (mvalid + NS) * W_SLAB);
                for (int j = 0; j < mvalid; j++) {
                    const int sl = slab0 + j, t = sl / p.kchunks, kc = sl - t * p.kchunks;
                    tma_load_5d(sa + j * W_SLAB, &tmA, &full[s], kc * KS, x0 + p.dx[t], y0 + p.dy[t], img0, g);
                }
#pragma unroll
                for (int j = 0; j < NS; j++) {
                    int c = n0 + j * KS, xx = x0, yy = y0;
                    if (p.cq) {
                        const int ph = c / p.cq;
                        c -= ph * p.cq;
                        xx = 2 * xx + (ph & 1);
                        yy = 2 * yy + (ph >> 1);
                    }
                    tma_load_5d(sg + j * W_SLAB, &tmG, &full[s], c, xx, yy, img0, g);
                }
            }
        } else if (warp == 1) {
            // ---------------- MMA issuer: D[128 x NT] += A_slabs^T[128 x 8 px] * Gd_slabs[8 px x NT], 8 times per stage.
            // Warp-uniform loop, one elected lane issues (descriptors stay in uniform registers).
            const uint32_t idesc = idesc_tf32_mn(NT);
            const uint32_t smem_base = smem_u32(smem);
            for (int it = 0; it < nsteps; it++) {
                const int s = it % STAGES, round = it / STAGES;
                mbar_wait(&full[s], round & 1);
                tc_fence_after();
                const uint32_t sa = smem_base + s * SM::STAGE, sg = sa + 4 * W_SLAB;
                const uint64_t da = smem_desc_mn_sw128(sa, W_SLAB), dg = smem_desc_mn_sw128(sg, W_SLAB);
#pragma unroll
                for (int k = 0; k < WPB / 8; k++)     // 8 pixels = 1024 bytes per MMA: +64 in the (>>4) start-address field
                    if (elect_one()) tc_mma_tf32(tmem, da + 64 * k, dg + 64 * k, idesc, (it | k) ? 1u : 0u);
                if (elect_one()) tc_commit(&empty[s]);
                __syncwarp();
            }
            if (elect_one()) tc_commit(accum);
            __syncwarp();
        }
        __syncwarp();
        // ---------------- epilogue: lane = input channel within the slab (contiguous in dW), column = output channel
        mbar_wait(accum, 0);
        tc_fence_after();
        if (warp < mvalid) {
            const int sl = slab0 + warp, t = sl / p.kchunks, kc = sl - t * p.kchunks;
            float *dst = p.dW + g * p.dw_gs + ((long long)t * p.N + n0) * p.Kt + kc * KS + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                float v[32];
                tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
                for (int j = 0; j < 32; j++) atomicAdd(dst + (long long)(c0 + j) * p.Kt, v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NT) : "memory");
}

bool tile_geometry64(int H, int W, int &bw, int &bh, int &bn) { return tile_geometry_n(H, W, WPB, bw, bh, bn); }

template <int NT, int STAGES>
int launch_wg(const CUtensorMap &tmA, const CUtensorMap &tmG, const WgParams &wp, dim3 grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_wgrad_tc<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<NT, STAGES>::BYTES));
        attr = true;
    }
    k_wgrad_tc<NT, STAGES><<<grid, 128, WgSmem<NT, STAGES>::BYTES, st>>>(tmA, tmG, wp);
    VV_CKL();
    return 0;
}

}  // namespace

bool vv_wgrad_tc_supported(const VvWGrad &p) {
    int bw, bh, bn;
    if (!encode_fn()) return false;
    if (p.Kt % KS || p.N % 32 || p.N < 32) return false;
    if (p.lda % 4 || p.a_coff % 4 || p.ldg % 4 || p.g_coff % 4 || p.a_gs % 4 || p.g_gs % 4) return false;
    if (((uintptr_t)p.A) % 16 || ((uintptr_t)p.Gd) % 16) return false;
    if (p.g_s2d && ((p.N / 4) % KS)) return false;
    if (!tile_geometry64(p.H, p.W, bw, bh, bn)) return false;
    return true;
}

int vv_launch_wgrad_tc(const VvWGrad &p, cudaStream_t st) {
    VV_REQUIRE(vv_wgrad_tc_supported(p), "wgrad_tc: unsupported shape (Kt=%d N=%d H=%d W=%d)", p.Kt, p.N, p.H, p.W);
    EncodeTiledFn enc = encode_fn();
    WgParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.B = p.B; wp.H = p.H; wp.W = p.W; wp.G = p.G;
    tile_geometry64(p.H, p.W, wp.bw, wp.bh, wp.bn);
    wp.tiles_x = p.W / wp.bw; wp.tiles_y = p.H / wp.bh; wp.tiles_n = (p.B + wp.bn - 1) / wp.bn;
    wp.ntaps = p.taps.n; wp.kchunks = p.Kt / KS; wp.nslabs = wp.ntaps * wp.kchunks;
    for (int t = 0; t < 9; t++) { wp.dy[t] = p.taps.dy[t]; wp.dx[t] = p.taps.dx[t]; }
    wp.N = p.N; wp.Kt = p.Kt; wp.cq = p.g_s2d ? p.N / 4 : 0;
    wp.dW = p.dW; wp.dw_gs = p.dw_gs;

    const CUtensorMapDataType dt = tmap_dtype();
    alignas(64) CUtensorMap tmA, tmG;
    {
        cuuint64_t dims[5] = {(cuuint64_t)p.Kt, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B, (cuuint64_t)p.G};
        cuuint64_t strides[4] = {(cuuint64_t)p.lda * 4, (cuuint64_t)p.W * p.lda * 4, (cuuint64_t)p.H * p.W * p.lda * 4,
                                 (cuuint64_t)(p.G > 1 ? p.a_gs : (long long)p.B * p.H * p.W * p.lda) * 4};
        cuuint32_t box[5] = {KS, (cuuint32_t)wp.bw, (cuuint32_t)wp.bh, (cuuint32_t)wp.bn, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmA, dt, 5, (void *)(p.A + p.a_coff), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "wgrad_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        const int sc = p.g_s2d ? 2 : 1;
        const cuuint64_t C = p.g_s2d ? p.N / 4 : p.N;
        cuuint64_t dims[5] = {C, (cuuint64_t)sc * p.W, (cuuint64_t)sc * p.H, (cuuint64_t)p.B, (cuuint64_t)p.G};
        cuuint64_t strides[4] = {(cuuint64_t)p.ldg * 4, (cuuint64_t)sc * p.W * p.ldg * 4, (cuuint64_t)sc * p.H * sc * p.W * p.ldg * 4,
                                 (cuuint64_t)(p.G > 1 ? p.g_gs : (long long)p.B * sc * p.H * sc * p.W * p.ldg) * 4};
        cuuint32_t box[5] = {KS, (cuuint32_t)(sc * wp.bw), (cuuint32_t)(sc * wp.bh), (cuuint32_t)wp.bn, 1};
        cuuint32_t estr[5] = {1, (cuuint32_t)sc, (cuuint32_t)sc, 1, 1};
        CUresult r = enc(&tmG, dt, 5, (void *)(p.Gd + p.g_coff), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "wgrad_tc: cuTensorMapEncodeTiled(Gd) failed with %d", (int)r);
    }
    const int nt_tile = p.N % 128 == 0 ? 128 : (p.N % 64 == 0 ? 64 : 32);
    const int mgroups = (wp.nslabs + 3) / 4;
    const int out_tiles = mgroups * (p.N / nt_tile) * p.G;
    const int tiles_total = wp.tiles_x * wp.tiles_y * wp.tiles_n;
    int splits = (2 * 148 + out_tiles - 1) / out_tiles;
    if (splits > tiles_total) splits = tiles_total;
    if (splits < 1) splits = 1;
    wp.tiles_per_split = (tiles_total + splits - 1) / splits;
    splits = (tiles_total + wp.tiles_per_split - 1) / wp.tiles_per_split;
    dim3 grid(mgroups * (p.N / nt_tile), splits, p.G);
    if (nt_tile == 128) return launch_wg<128, 3>(tmA, tmG, wp, grid, st);
    if (nt_tile == 64) return launch_wg<64, 4>(tmA, tmG, wp, grid, st);
    return launch_wg<32, 4>(tmA, tmG, wp, grid, st);
}
