// FlowNetC cost volume (kernel 1, stride1 1, stride2 2, 21x21 displacements; FlowNetC.py:24-30), TMA-fed version.
//
// out[n, tj*21+ti, y, x] = 1/C * sum_c in1[n,c,y+o,x+o] * in2[n,c,y+o+2tj,x+o+2ti]     (o = max_displacement - pad, zero padded)
//
// The stride-2 displacement grid means an output pixel only ever meets in2 columns of ITS OWN parity.  A light pre-pass
// splits both feature maps by column parity ([B][parity][C][H][W/2], the only scratch: the reference instead built two
// zero-padded NHWC copies and zero-filled three tensors, correlation_cuda.c:36-42).  The main kernel then streams, per
// chunk of 8 channels, four TMA boxes into shared memory -- for each parity one 32-wide row of in1 and 21 rows (every
// second image row, a TMA element stride) x 56 columns of in2, halo zero-filled by TMA -- through a 3-stage mbarrier ring
// fed by a dedicated producer warp.  Compute threads are (tj, parity, quad): 4 output pixels x 21 ti = 84 accumulators,
// operands fetched with 128-bit shared-memory loads (7 + 1 per channel for 84 FMAs).  Requires pad == max_displacement
// (FlowNetC) so that box starts fall on 16-byte boundaries; other parameters use the kernels in flow_ops.cu.
#include "tc_common.cuh"

namespace {

constexpr int CD = 21, CDR = 10, CTX = 64, CCC = 8;        // displacement grid, x tile, channels per stage
constexpr int C2H = 56;                                     // in2 half-columns per row: 32 + 2*10, + 4 because a TMA box must start
                                                            // on a 16-byte boundary in its innermost dimension (start h0 - 12, not h0 - 10)
constexpr int C_F1 = CCC * 32, C_F2 = CCC * CD * C2H;       // floats per parity per stage
constexpr int C_STAGE = 2 * (C_F1 + C_F2);                  // floats per stage (both parities)
constexpr int C_STAGES = 3;
constexpr int C_COMPUTE = 16 * CD;                          // 336 compute threads
constexpr int C_THREADS = 384;                              // 11 compute warps (352, 336 used) + 1 producer warp

struct CorrTmaParams {
    int B, C, H, W, W2p, OH, OW, OC, off;
};

// ---- pre-pass: [B][C][H][W] -> [B][2][C][H][W2p], column x -> plane x&1, index x>>1; padding columns zeroed
__global__ void k_parity_split(const float *__restrict__ in, float *__restrict__ out, int C, int H, int W, int W2p) {
    const long long rows = (long long)C * H;                 // per image
    const int b = blockIdx.y;
    for (long long r = blockIdx.x * (long long)blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y) {
        const float *src = in + ((long long)b * rows + r) * W;
        float *d0 = out + (((long long)b * 2 + 0) * rows + r) * W2p, *d1 = out + (((long long)b * 2 + 1) * rows + r) * W2p;
        for (int h = threadIdx.x; h < W2p; h += blockDim.x) {
            const int x = 2 * h;
            d0[h] = x < W ? src[x] : 0.f;
            d1[h] = x + 1 < W ? src[x + 1] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(C_THREADS, 1) k_corr_fwd_tma(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
                                                               float *__restrict__ out, const CorrTmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    // aligned by an offset, not by an integer round trip: the pointer keeps its shared address space and the inner loop's operand
    // fetches compile to LDS rather than generic LD
    float *sm = (float *)(smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u));
    uint64_t *full = (uint64_t *)(sm + C_STAGES * C_STAGE);
    uint64_t *empty = full + C_STAGES;
    const int x0 = blockIdx.x * CTX, y = blockIdx.y, n = blockIdx.z;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int y1 = y + p.off;
    if (t == 0) {
        for (int s = 0; s < C_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 11); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm2) : "memory");
    }
    __syncthreads();
    const int nchunk = (p.C + CCC - 1) / CCC;

    if (warp == 11) {
        if (lane == 0) {
            // ---------------- TMA producer: per parity one in1 row box and one in2 (21 rows x 52) box per channel chunk
            for (int ch = 0; ch < nchunk; ch++) {
                const int s = ch % C_STAGES, round = ch / C_STAGES;
                if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                float *st = sm + s * C_STAGE;
                mbar_expect_tx(&full[s], C_STAGE * 4);
#pragma unroll
                for (int pp = 0; pp < 2; pp++) {
                    const int xb = x0 + p.off + pp;                 // first in1 column of this thread parity
                    const int gp = xb & 1;                          // parity plane it lives in
                    const int h0 = (xb - gp) / 2;                   // (xb - gp) is even: exact
                    tma_load_4d(st + pp * C_F1, &tm1, &full[s], h0, y1, ch * CCC, n * 2 + gp);
                    tma_load_4d(st + 2 * C_F1 + pp * C_F2, &tm2, &full[s], h0 - CDR - 2, y1 - 2 * CDR, ch * CCC, n * 2 + gp);
                }
            }
        }
    } else {
        const bool active = t < C_COMPUTE;
        const int tj = active ? t / 16 : 0, xg = t % 16, pp = xg >> 3, q = xg & 7;
        float acc[4][CD];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int i = 0; i < CD; i++) acc[k][i] = 0.f;
        for (int ch = 0; ch < nchunk; ch++) {
            const int s = ch % C_STAGES;
            mbar_wait(&full[s], (ch / C_STAGES) & 1);
            const float *st = sm + s * C_STAGE;
            // the window a thread needs starts 2 floats into a 16-byte group (the box starts at h0 - 12, the window at h0 - 10): it reads
            // the 28 floats from the aligned group on with seven 16-byte loads -- the eight q-lanes of a row then cover 128 contiguous
            // bytes per load (one wavefront per row) where twelve 8-byte loads at a 16-byte stride used half of every wavefront
            const float *f1 = st + pp * C_F1 + 4 * q, *f2 = st + 2 * C_F1 + pp * C_F2 + tj * C2H + 4 * q;
            if (active) {
#pragma unroll
                for (int c = 0; c < CCC; c++) {
                    const float4 av = *reinterpret_cast<const float4 *>(f1 + c * 32);
                    const float *row = f2 + c * CD * C2H;
                    float bv[28];
#pragma unroll
                    for (int m = 0; m < 7; m++) {
                        const float4 v = *reinterpret_cast<const float4 *>(row + 4 * m);
                        bv[4 * m] = v.x; bv[4 * m + 1] = v.y; bv[4 * m + 2] = v.z; bv[4 * m + 3] = v.w;
                    }
                    const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int i = 0; i < CD; i++) acc[k][i] = fmaf(aa[k], bv[2 + k + i], acc[k][i]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // Epilogue through shared memory (the ring is idle now: every stage was consumed): a thread's 84 results belong to 21
        // displacement planes x 4 scattered columns, so written directly they are 84 four-byte stores per thread; staged as
        // [441 planes][64 columns] they leave as 16-byte stores covering whole 256-byte output rows.
        asm volatile("bar.sync 1, 352;" ::: "memory");          // the eleven compute warps: nobody still reads the last stage
        constexpr int OUT_LD = CTX + 4;                         // 68 floats: 16-byte aligned rows, the two tj of a warp on different banks
        float *so = sm;
        if (active) {
            const float inv = 1.f / (float)p.C;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int xl = pp + 2 * (4 * q + k);
#pragma unroll
                for (int i = 0; i < CD; i++) so[(tj * CD + i) * OUT_LD + xl] = acc[k][i] * inv;
            }
        }
        asm volatile("bar.sync 1, 352;" ::: "memory");
        {
            const long long plane = (long long)p.OH * p.OW;
            float *o = out + (long long)n * p.OC * plane + (long long)y * p.OW + x0;
            const bool vec = (p.OW % 4 == 0) && x0 + CTX <= p.OW;      // whole tile inside the row and 16-byte aligned
            for (int e = t; e < CD * CD * (CTX / 4); e += 352) {
                const int d = e / (CTX / 4), c4 = (e - d * (CTX / 4)) * 4;
                const float4 v = *reinterpret_cast<const float4 *>(so + d * OUT_LD + c4);
                float *dst = o + d * plane + c4;
                if (vec) {
                    *reinterpret_cast<float4 *>(dst) = v;
                } else {
                    if (x0 + c4 < p.OW) dst[0] = v.x;
                    if (x0 + c4 + 1 < p.OW) dst[1] = v.y;
                    if (x0 + c4 + 2 < p.OW) dst[2] = v.z;
                    if (x0 + c4 + 3 < p.OW) dst[3] = v.w;
                }
            }
        }
    }
}

}  // namespace

// bytes of scratch the TMA path needs (0: shape not handled by it)
long long vv_corr_tma_workspace(int batch, int channels, int h, int w) {
    if (!encode_fn()) return 0;
    const int w2p = ((w + 1) / 2 + 3) / 4 * 4;
    return 2LL * batch * 2 * channels * h * w2p * (long long)sizeof(float);
}

int vv_launch_corr_tma(const float *in1, const float *in2, float *out, int batch, int channels, int h, int w, int pad, int md, int oc, int oh,
                       int ow, void *workspace, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    VV_REQUIRE(enc, "correlation: cuTensorMapEncodeTiled unavailable");
    const int w2p = ((w + 1) / 2 + 3) / 4 * 4;
    float *s1 = (float *)workspace, *s2 = s1 + (long long)batch * 2 * channels * h * w2p;
    VV_REQUIRE(((uintptr_t)workspace) % 16 == 0, "correlation: workspace must be 16-byte aligned");
    {
        dim3 blk(32, 8);
        int gx = vv_cdiv((long long)channels * h, 8);
        if (gx > 148 * 8) gx = 148 * 8;
        k_parity_split<<<dim3(gx, batch), blk, 0, st>>>(in1, s1, channels, h, w, w2p);
        VV_CKL();
        k_parity_split<<<dim3(gx, batch), blk, 0, st>>>(in2, s2, channels, h, w, w2p);
        VV_CKL();
    }
    alignas(64) CUtensorMap tm1, tm2;
    cuuint64_t dims[4] = {(cuuint64_t)w2p, (cuuint64_t)h, (cuuint64_t)channels, (cuuint64_t)batch * 2};
    cuuint64_t strides[3] = {(cuuint64_t)w2p * 4, (cuuint64_t)h * w2p * 4, (cuuint64_t)channels * h * w2p * 4};
    {
        cuuint32_t box[4] = {32, 1, CCC, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s1, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "correlation: cuTensorMapEncodeTiled(in1) failed with %d", (int)r);
    }
    {
        cuuint32_t box[4] = {C2H, 2 * CD, CCC, 1};          // 21 rows at element stride 2
        cuuint32_t estr[4] = {1, 2, 1, 1};
        CUresult r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, s2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        VV_REQUIRE(r == CUDA_SUCCESS, "correlation: cuTensorMapEncodeTiled(in2) failed with %d", (int)r);
    }
    CorrTmaParams p;
    p.B = batch; p.C = channels; p.H = h; p.W = w; p.W2p = w2p; p.OH = oh; p.OW = ow; p.OC = oc; p.off = md - pad;
    const int smem = C_STAGES * C_STAGE * 4 + 128 + 64;
    static bool attr = false;
    if (!attr) {
        VV_CK(cudaFuncSetAttribute(k_corr_fwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    k_corr_fwd_tma<<<dim3(vv_cdiv(ow, CTX), oh, batch), C_THREADS, smem, st>>>(tm1, tm2, out, p);
    VV_CKL();
    return 0;
}
