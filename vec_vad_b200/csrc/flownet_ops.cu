// Kernels of the FlowNet2 inference graph (reference: FlowNet2_src/models/flownet2.py:65-149 and
// components/FlowNetC.py, FlowNetS.py, FlowNetSD.py, FlowNetFusion.py, misc.py:6-45) that are not the three native ops
// (correlation / resample2d / channelnorm live in corr_tma.cu / flow_ops.cu).
//
// Layout: NCHW fp32 like the reference, but every tensor argument is a CHANNEL SLICE of a possibly larger buffer: a pointer to
// the slice's first channel plus the batch stride of the buffer it lives in.  torch.cat([...], 1) of the reference is therefore
// never executed: the producers write their channels straight into the concat buffer.
//
//   k_fn_conv<TN>   Conv2d (k in {1,3,5,7}, stride 1/2, pad (k-1)/2) and ConvTranspose2d(4, 2, 1) as ONE gather-GEMM:
//                   out[b, co, gy*o_mul + o_off, ..] = bias[co] + sum_{ci, t} in[b, ci, gy*i_mul + ty(t), gx*i_mul + tx(t)] * w[co][ci*ntaps + t]
//                   (+ LeakyReLU(0.1), misc.py:25-26).  The transposed conv runs as its four output-parity phases, each a 2x2-tap
//                   gather with the phase's weights re-laid out once at load time, all four in one launch.  fp32 FMA tiles:
//                   128 pixels x TN channels per CTA (TN = 16 ... 128, picked per layer shape by plan_conv), K staged through shared
//                   memory in chunks of 16 with a register prefetch of the next chunk.
//                   Arithmetic is exact fp32 (the reference runs these layers in fp32 cuDNN); summation order differs.
//   k_fn_mean / k_fn_normalize   per-(image, colour) mean over both frames, (x - mean) / rgb_max, frames concatenated on channels
//                   (flownet2.py:66-72).
//   k_fn_upsample4  nn.Upsample(scale_factor=4, 'bilinear' | 'nearest') times a constant (flownet2.py:28,34,41-42,76,90,105,122).
//   k_fn_scale_copy slice -> slice copy with a scale and an optional LeakyReLU (concat members that are not conv outputs,
//                   the correlation's activation FlowNetC.py:33,91).
#include "common.h"

namespace {

constexpr int FN_KC = 16;                             // granularity of a split of the contraction (every tile's chunk divides it)

struct FnConv {
    const float *in;   long long in_bs;  int Cin, IH, IW;
    const float *w;                                      // [Co][K], K = Cin * ntaps
    const float *bias;                                   // nullable
    float *out;        long long out_bs; int Co, OH, OW;
    int GH, GW;                                          // pixel grid of this launch
    int o_mul, o_off_y, o_off_x;                         // output pixel = g * o_mul + o_off
    int i_mul;                                           // input pixel  = g * i_mul + tap offset
    // taps: t = ti * kw + tj, ti < kh, tj < kw, at input offset (ty0 + dty * ti, tx0 + dtx * tj).  A k x k convolution is
    // (kh, kw, ty0, dty, tx0, dtx) = (k, k, -pad, 1, -pad, 1); an output-parity phase of the transposed conv is (2, 2, 0 or 1, -1, 0 or 1, -1)
    int ntaps, kh, kw, ty0, dty, tx0, dtx;
    unsigned m_taps, m_kw;                               // ceil(2^32 / ntaps), ceil(2^32 / kw): n / d = umulhi(n, m) for n < 2^16
    int B, act;                                          // act: LeakyReLU(0.1)
    int ksplit, kper;                                    // split-K: blockIdx.z handles k in [z * kper, (z + 1) * kper); kper % 16 == 0
    float *partial;                                      // ksplit > 1: raw partial sums [ksplit][Co][M], finished by k_fn_conv_finish
    int w_vec, out_vec, part_vec;                        // 16-byte weight loads / output stores / partial-sum stores are legal
    // phases == 4: the four output-parity phases of ConvTranspose2d(4, 2, 1) in ONE launch.  blockIdx.z = phase * ksplit + split;
    // phase (py, px) = (ph >> 1, ph & 1) reads the weights w + ph * Co * K, starts its taps at (ty0, tx0) = (py, px) and writes the
    // output pixels (2 gy + py, 2 gx + px)
    int phases;
};

// One CTA of 256 threads computes TM pixels x TN channels; a thread owns 8 pixels (two groups of 4: 16-byte shared-memory reads and
// 16-byte stores) x CN = TM * TN / 2048 channels.  The shapes in use (plan_conv picks one per layer):
//   128 x 128, 256 x 64   8 x 8 accumulators: 64 FMAs per four 16-byte operand reads; two CTAs per SM
//   128 x 64, 256 x 32    8 x 4
//   512 x 16 (KC = 8)     8 x 4 for the 16-channel layers, where a 128-pixel tile would leave 8 FMAs per three operand reads
//   128 x 32, 128 x 16    8 x 2, 8 x 1: small pixel grids of the narrow layers
template <int TM, int TN, int KC>
__global__ void __launch_bounds__(256, (TM * TN >= 16384 || TM >= 512) ? 2 : (TM * TN >= 8192 ? 3 : 4)) k_fn_conv(const FnConv p) {
    constexpr int CN = TM * TN / 2048;                   // output channels per thread
    constexpr int TMT = TM / 8;                          // threads along the pixels
    constexpr int NPIX = TM > 256 ? TM / 256 : 1;        // loader: pixels per thread
    constexpr int PIXT = TM / NPIX;                      // loader: pixels covered by one slice of threads
    constexpr int KPT = KC / (256 / PIXT);               // loader: consecutive k per thread and pixel
    constexpr int BPT = TN * KC >= 256 ? TN * KC / 256 : 1;   // weight elements per (loading) thread
    static_assert(CN == 1 || CN == 2 || CN == 4 || CN == 8, "tile shape");
    static_assert(KPT >= 1 && KPT * (256 / PIXT) == KC && FN_KC % KC == 0, "chunk shape");
    __shared__ __align__(16) float As[2][KC][TM + 4];
    __shared__ __align__(16) float Bs[2][KC][TN + 4];
    vv_pdl_wait();
    const int tid = threadIdx.x;
    const int Kall = p.Cin * p.ntaps;
    const int ph = p.phases > 1 ? blockIdx.z / p.ksplit : 0, zsplit = blockIdx.z - ph * p.ksplit;
    const int py = p.phases > 1 ? ph >> 1 : 0, px = p.phases > 1 ? ph & 1 : 0;
    const int kbeg = zsplit * p.kper, K = min(Kall, kbeg + p.kper);        // this CTA's share of the contraction
    const long long M = (long long)p.B * p.GH * p.GW;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    // ---- loader roles.  A: thread owns NPIX pixels and KPT consecutive k of the chunk; B: a channel and BPT consecutive k
    const int a_k0 = (tid / PIXT) * KPT;
    const float *a_base[NPIX];                           // the pixel's own input position (tap offset (0, 0), channel 0)
    int iy0[NPIX], ix0[NPIX];                            // a pixel beyond M fails every bounds check
#pragma unroll
    for (int q = 0; q < NPIX; q++) {
        const long long am = m0 + (tid % PIXT) + q * PIXT;
        a_base[q] = p.in; iy0[q] = 1 << 21; ix0[q] = 0;
        if (am < M) {
            const int ab = (int)(am / ((long long)p.GH * p.GW));
            const int r = (int)(am - (long long)ab * p.GH * p.GW);
            const int agy = r / p.GW, agx = r - agy * p.GW;
            iy0[q] = agy * p.i_mul; ix0[q] = agx * p.i_mul;
            a_base[q] = p.in + (long long)ab * p.in_bs + iy0[q] * p.IW + ix0[q];
        }
    }
    const int IHW = p.IH * p.IW;
    const int row_adv = p.dty * p.IW - p.kw * p.dtx, ci_adv = IHW - p.kh * p.dty * p.IW;   // offset steps when tj / ti wrap
    const bool b_act = TN * KC >= 256 || tid < TN * KC;
    const int b_n = (tid * BPT) / KC, b_k0 = (tid * BPT) % KC;
    const bool b_ok = b_act && n0 + b_n < p.Co;
    const float *b_base = p.w + ((long long)ph * p.Co + n0 + b_n) * Kall;

    float ra[NPIX][KPT], rb[BPT];
    auto load = [&](int k0) {
        // (ci, ti, tj) of the first k by two reciprocal multiplies, then the consecutive k walk taps / rows / channels with a running
        // tap offset (dy, dx) and a running element offset: no table, no division; the walk is shared by the thread's pixels
        int kg = k0 + a_k0;
        const int ci = p.ntaps == 1 ? kg : (int)__umulhi((unsigned)kg, p.m_taps);
        const int t = kg - ci * p.ntaps;
        int ti = p.kw == 1 ? t : (int)__umulhi((unsigned)t, p.m_kw);
        int tj = t - ti * p.kw;
        int dy = p.ty0 + py + p.dty * ti, dx = p.tx0 + px + p.dtx * tj;
        int off = ci * IHW + dy * p.IW + dx;
#pragma unroll
        for (int j = 0; j < KPT; j++, kg++) {
#pragma unroll
            for (int q = 0; q < NPIX; q++) {
                const bool ok = kg < K && (unsigned)(iy0[q] + dy) < (unsigned)p.IH && (unsigned)(ix0[q] + dx) < (unsigned)p.IW;
                ra[q][j] = ok ? __ldg(a_base[q] + off) : 0.f;
            }
            tj++; dx += p.dtx; off += p.dtx;
            if (tj == p.kw) {
                tj = 0; dx -= p.kw * p.dtx; dy += p.dty; off += row_adv; ti++;
                if (ti == p.kh) { ti = 0; dy -= p.kh * p.dty; off += ci_adv; }
            }
        }
        if constexpr (BPT >= 4) {
            if (p.w_vec) {                                // K and every split boundary are multiples of 4: a group is all in or all out
#pragma unroll
                for (int j = 0; j < BPT; j += 4) {
                    const int kb = k0 + b_k0 + j;
                    const float4 v = (b_ok && kb < K) ? __ldg(reinterpret_cast<const float4 *>(b_base + kb)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    rb[j] = v.x; rb[j + 1] = v.y; rb[j + 2] = v.z; rb[j + 3] = v.w;
                }
                return;
            }
        }
#pragma unroll
        for (int j = 0; j < BPT; j++) {
            const int kb = k0 + b_k0 + j;
            rb[j] = (b_ok && kb < K) ? __ldg(b_base + kb) : 0.f;
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int q = 0; q < NPIX; q++)
#pragma unroll
            for (int j = 0; j < KPT; j++) As[buf][a_k0 + j][(tid % PIXT) + q * PIXT] = ra[q][j];
        if (b_act) {
#pragma unroll
            for (int j = 0; j < BPT; j++) Bs[buf][b_k0 + j][b_n] = rb[j];
        }
    };
    // ---- compute roles: TM / 8 x (256 / (TM / 8)) threads, thread (tm, tn) owns pixels 4 tm .. 4 tm + 3 and TM / 2 + 4 tm .. + 3 (two
    // 16-byte shared-memory reads per k instead of eight scalar ones) and channels tn * CN + j
    const int tm = tid % TMT, tn = tid / TMT;
    float acc[8][CN];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < CN; j++) acc[i][j] = 0.f;

    load(kbeg);
    stage(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < K; k0 += KC) {
        const bool more = k0 + KC < K;
        if (more) load(k0 + KC);                          // global loads of the next chunk in flight during the FMAs
#pragma unroll
        for (int k = 0; k < KC; k++) {
            float a[8], b[CN];
            *reinterpret_cast<float4 *>(a) = *reinterpret_cast<const float4 *>(&As[buf][k][4 * tm]);
            *reinterpret_cast<float4 *>(a + 4) = *reinterpret_cast<const float4 *>(&As[buf][k][TM / 2 + 4 * tm]);
            if constexpr (CN == 8) {
                *reinterpret_cast<float4 *>(b) = *reinterpret_cast<const float4 *>(&Bs[buf][k][8 * tn]);
                *reinterpret_cast<float4 *>(b + 4) = *reinterpret_cast<const float4 *>(&Bs[buf][k][8 * tn + 4]);
            } else if constexpr (CN == 4) *reinterpret_cast<float4 *>(b) = *reinterpret_cast<const float4 *>(&Bs[buf][k][4 * tn]);
            else if constexpr (CN == 2) *reinterpret_cast<float2 *>(b) = *reinterpret_cast<const float2 *>(&Bs[buf][k][2 * tn]);
            else b[0] = Bs[buf][k][tn];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < CN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            stage(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    // ---- epilogue: bias, LeakyReLU, NCHW store; split-K: raw partials.  Whole groups of 4 pixels leave as one 16-byte store when
    // the launch allows it (the 4 pixels then share an output row); otherwise consecutive lanes cover consecutive pixel groups.
    const long long OHW = (long long)p.OH * p.OW;
    const bool split = p.ksplit > 1;
    if (split ? p.part_vec : p.out_vec) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const long long m = m0 + (TM / 2) * g + 4 * tm;
            if (m >= M) continue;                         // M % 4 == 0: the group is whole
            float *o;
            if (split) o = p.partial + (long long)blockIdx.z * p.Co * M + m;
            else {
                const int b = (int)(m / ((long long)p.GH * p.GW));
                const int r = (int)(m - (long long)b * p.GH * p.GW);
                o = p.out + (long long)b * p.out_bs + r;  // o_mul == 1, no offset: the output plane is the pixel grid
            }
#pragma unroll
            for (int j = 0; j < CN; j++) {
                const int co = n0 + tn * CN + j;
                if (co >= p.Co) continue;
                float4 v = make_float4(acc[4 * g][j], acc[4 * g + 1][j], acc[4 * g + 2][j], acc[4 * g + 3][j]);
                if (!split) {
                    const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
                    v.x += bv; v.y += bv; v.z += bv; v.w += bv;
                    if (p.act) {
                        v.x = v.x > 0.f ? v.x : 0.1f * v.x; v.y = v.y > 0.f ? v.y : 0.1f * v.y;
                        v.z = v.z > 0.f ? v.z : 0.1f * v.z; v.w = v.w > 0.f ? v.w : 0.1f * v.w;
                    }
                }
                *reinterpret_cast<float4 *>(o + (long long)co * (split ? M : OHW)) = v;
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long m = m0 + (i < 4 ? 4 * tm + i : TM / 2 + 4 * tm + i - 4);
        if (m >= M) continue;
        if (split) {
#pragma unroll
            for (int j = 0; j < CN; j++) {
                const int co = n0 + tn * CN + j;
                if (co < p.Co) p.partial[((long long)blockIdx.z * p.Co + co) * M + m] = acc[i][j];
            }
            continue;
        }
        const int b = (int)(m / ((long long)p.GH * p.GW));
        const int r = (int)(m - (long long)b * p.GH * p.GW);
        const int gy = r / p.GW, gx = r - gy * p.GW;
        float *o = p.out + (long long)b * p.out_bs + (long long)(gy * p.o_mul + p.o_off_y + py) * p.OW + (gx * p.o_mul + p.o_off_x + px);
#pragma unroll
        for (int j = 0; j < CN; j++) {
            const int co = n0 + tn * CN + j;
            if (co < p.Co) {
                float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
                if (p.act) v = v > 0.f ? v : 0.1f * v;
                o[co * OHW] = v;
            }
        }
    }
}

// out = act(bias + sum over the splits of the partial sums)
__global__ void k_fn_conv_finish(const FnConv p) {
    vv_pdl_wait();
    const long long M = (long long)p.B * p.GH * p.GW, per_phase = M * p.Co, total = per_phase * p.phases;
    const long long OHW = (long long)p.OH * p.OW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ph = (int)(i / per_phase);
        const long long r0 = i - ph * per_phase;
        const int co = (int)(r0 / M);
        const long long m = r0 - (long long)co * M;
        float v = p.bias ? __ldg(p.bias + co) : 0.f;
        for (int z = 0; z < p.ksplit; z++) v += p.partial[((long long)(ph * p.ksplit + z) * p.Co + co) * M + m];
        if (p.act) v = v > 0.f ? v : 0.1f * v;
        const int b = (int)(m / ((long long)p.GH * p.GW));
        const int r = (int)(m - (long long)b * p.GH * p.GW);
        const int gy = r / p.GW, gx = r - gy * p.GW;
        const int py = p.phases > 1 ? ph >> 1 : 0, px = p.phases > 1 ? ph & 1 : 0;
        p.out[(long long)b * p.out_bs + co * OHW + (long long)(gy * p.o_mul + p.o_off_y + py) * p.OW + (gx * p.o_mul + p.o_off_x + px)] = v;
    }
}

// The same, four consecutive pixels per thread (M % 4 == 0, partial sums 16-byte aligned): 16-byte reads of the partial sums, several
// splits' loads in flight, one 16-byte store where the output plane is the pixel grid.  Same order of additions as the scalar form.
__global__ void k_fn_conv_finish4(const FnConv p) {
    vv_pdl_wait();
    const long long M = (long long)p.B * p.GH * p.GW, M4 = M / 4, per_phase = M4 * p.Co, total = per_phase * p.phases;
    const long long OHW = (long long)p.OH * p.OW, GHW = (long long)p.GH * p.GW;
    const long long zs = (long long)p.Co * M4;          // float4 stride between two splits
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ph = (int)(i / per_phase);
        const long long r0 = i - ph * per_phase;
        const int co = (int)(r0 / M4);
        const long long m = (r0 - (long long)co * M4) * 4;
        const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
        float4 v = make_float4(bv, bv, bv, bv);
        const float4 *src = reinterpret_cast<const float4 *>(p.partial + ((long long)ph * p.ksplit * p.Co + co) * M + m);
#pragma unroll 4
        for (int z = 0; z < p.ksplit; z++) {
            const float4 t = src[z * zs];
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        if (p.act) {
            v.x = v.x > 0.f ? v.x : 0.1f * v.x; v.y = v.y > 0.f ? v.y : 0.1f * v.y;
            v.z = v.z > 0.f ? v.z : 0.1f * v.z; v.w = v.w > 0.f ? v.w : 0.1f * v.w;
        }
        if (p.out_vec) {                                  // o_mul == 1, no offsets, GW % 4 == 0: the four pixels are contiguous in the output
            const int b = (int)(m / GHW);
            *reinterpret_cast<float4 *>(p.out + (long long)b * p.out_bs + co * OHW + (m - (long long)b * GHW)) = v;
        } else {
            const int py = p.phases > 1 ? ph >> 1 : 0, px = p.phases > 1 ? ph & 1 : 0;
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const long long mq = m + q;
                const int b = (int)(mq / GHW);
                const int r = (int)(mq - (long long)b * GHW);
                const int gy = r / p.GW, gx = r - gy * p.GW;
                p.out[(long long)b * p.out_bs + co * OHW + (long long)(gy * p.o_mul + p.o_off_y + py) * p.OW + (gx * p.o_mul + p.o_off_x + px)] = e[q];
            }
        }
    }
}

// Tile shape and split of the contraction, chosen by a small cost model (all shapes static per layer, so the choice -- and with it
// the summation order -- is a function of the layer's shape alone):
//   cycles = CTAs on the busiest SM x FMAs per CTA / (128 lanes x the tile's measured FMA-pipe efficiency)
//            + for a split: the partial sums' trip through the scratch buffer and the finishing launch.
// The efficiencies are fitted to isolated-layer timings of every tile on eight FlowNet2 layer shapes
// (profiles/r02_fn_conv_tiles.txt).  Wide tiles amortise the gather (one input element feeds TN channels) and the shared-memory reads
// (64 FMAs per four 16-byte reads with 8 x 8 accumulators); narrow ones waste no columns on the 2- / 16- / 32-channel layers; short
// ones quantise better on small pixel grids.
struct FnTileInfo { int tm, tn, occ; double eff; };
const FnTileInfo FN_TILES[] = {{128, 16, 4, 0.18}, {128, 32, 4, 0.33}, {128, 64, 3, 0.48}, {128, 128, 2, 0.53},
                               {256, 64, 2, 0.49}, {256, 32, 3, 0.48}, {512, 16, 2, 0.36}};
constexpr int FN_NTILES = sizeof(FN_TILES) / sizeof(FN_TILES[0]);
struct FnPlan { int tile, ksplit; };
FnPlan plan_conv(const FnConv &p, bool have_scratch, long long scratch_floats) {
    const long long M = (long long)p.B * p.GH * p.GW;
    const int K = p.Cin * p.ntaps;
    static int force = -2;                                // VECVAD_FN_TILE=0..6 pins the tile shape (measurement knob)
    if (force < -1) { const char *e = getenv("VECVAD_FN_TILE"); force = e ? atoi(e) : -1; }
    static double split_scale = -1.0;                     // VECVAD_FN_SPLIT_COST scales the modelled cost of a split (measurement knob)
    if (split_scale < 0.0) { const char *e = getenv("VECVAD_FN_SPLIT_COST"); split_scale = e ? atof(e) : 1.0; }
    FnPlan best = {2, 1};
    double best_c = 1e300;
    for (int v = 0; v < FN_NTILES; v++) {
        const FnTileInfo &t = FN_TILES[v];
        if (force >= 0 ? v != force : (t.tn > 16 && p.Co <= t.tn / 2)) continue;   // a tile at most half full never wins
        const long long base = (long long)vv_cdiv(M, t.tm) * vv_cdiv(p.Co, t.tn) * p.phases;
        const int max_split = have_scratch ? (K / 64 < 32 ? K / 64 : 32) : 1;
        for (int ks = 1; ks <= (max_split < 1 ? 1 : max_split); ks++) {
            if (ks > 1 && (long long)ks * p.phases * p.Co * M > scratch_floats) break;
            const int kper = vv_cdiv(vv_cdiv(K, ks), FN_KC) * FN_KC;
            if (vv_cdiv(K, kper) != ks) continue;          // rounding kper up left the last split empty: same as a smaller ks
            const long long per_sm = (base * ks + 147) / 148;
            const double fill = per_sm >= t.occ ? 1.0 : 0.6 + 0.4 * (double)per_sm / t.occ;   // fewer resident warps hide less latency
            double c = (double)per_sm * t.tm * t.tn * kper / (128.0 * t.eff * fill) + 4000.0;  // + prologue / epilogue of a CTA
            if (ks > 1) c += split_scale * (8000.0 + 2.0 * ks * p.phases * p.Co * M * 4.0 / 1500.0);   // finishing launch + partials out and back
            if (c < best_c) { best_c = c; best.tile = v; best.ksplit = ks; }
        }
    }
    return best;
}

int launch_conv(FnConv p, float *scratch, long long scratch_floats, cudaStream_t st) {
    const long long M = (long long)p.B * p.GH * p.GW;
    const int K = p.Cin * p.ntaps;
    VV_REQUIRE(K < 65536, "fn_conv: contraction of %d terms too long", K);
    VV_REQUIRE((long long)p.Cin * p.IH * p.IW < (1LL << 31), "fn_conv: input image of %lld elements too large", (long long)p.Cin * p.IH * p.IW);
    p.kh = p.ntaps / p.kw;
    if (p.phases < 1) p.phases = 1;
    p.m_taps = p.ntaps > 1 ? 0xFFFFFFFFu / (unsigned)p.ntaps + 1u : 0u;
    p.m_kw = p.kw > 1 ? 0xFFFFFFFFu / (unsigned)p.kw + 1u : 0u;
    const FnPlan plan = plan_conv(p, scratch != nullptr, scratch_floats);
    const FnTileInfo &t = FN_TILES[plan.tile];
    const int ksplit = plan.ksplit;
    p.ksplit = ksplit;
    p.kper = vv_cdiv(vv_cdiv(K, ksplit), FN_KC) * FN_KC;
    p.partial = scratch;
    p.w_vec = K % 4 == 0 && ((uintptr_t)p.w) % 16 == 0;
    p.out_vec = p.o_mul == 1 && p.o_off_y == 0 && p.o_off_x == 0 && p.GW % 4 == 0 && p.OW == p.GW && p.OH == p.GH && ((uintptr_t)p.out) % 16 == 0 &&
                p.out_bs % 4 == 0;
    p.part_vec = M % 4 == 0 && ((uintptr_t)scratch) % 16 == 0;
    const dim3 grid(vv_cdiv(M, t.tm), vv_cdiv(p.Co, t.tn), ksplit * p.phases);
    cudaError_t e;
    switch (plan.tile) {
        case 0: e = vv_launch(k_fn_conv<128, 16, 16>, grid, dim3(256), 0, st, p); break;
        case 1: e = vv_launch(k_fn_conv<128, 32, 16>, grid, dim3(256), 0, st, p); break;
        case 2: e = vv_launch(k_fn_conv<128, 64, 16>, grid, dim3(256), 0, st, p); break;
        case 3: e = vv_launch(k_fn_conv<128, 128, 16>, grid, dim3(256), 0, st, p); break;
        case 4: e = vv_launch(k_fn_conv<256, 64, 16>, grid, dim3(256), 0, st, p); break;
        case 5: e = vv_launch(k_fn_conv<256, 32, 16>, grid, dim3(256), 0, st, p); break;
        default: e = vv_launch(k_fn_conv<512, 16, 8>, grid, dim3(256), 0, st, p); break;
    }
    VV_CK(e);
    VV_CKL();
    if (ksplit > 1) {
        const long long total = M * p.Co * p.phases / (p.part_vec ? 4 : 1);
        const dim3 fgrid((unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256));
        VV_CK(p.part_vec ? vv_launch(k_fn_conv_finish4, fgrid, dim3(256), 0, st, p) : vv_launch(k_fn_conv_finish, fgrid, dim3(256), 0, st, p));
        VV_CKL();
    }
    return 0;
}

// sums[b * 3 + c] += sum over this block's share of the 2 * H * W values of colour c of image pair b  (ims: [B,3,2,H,W])
__global__ void k_fn_mean(const float *__restrict__ ims, double *__restrict__ sums, long long per) {
    vv_pdl_wait();
    const int bc = blockIdx.y;
    const float *src = ims + (long long)bc * per;
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) s += (double)src[i];
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(sums + bc, red[0]);
}
// x[b, f*3 + c, y, x] = (ims[b, c, f, y, x] - mean[b, c]) / rgb_max
__global__ void k_fn_normalize(const float *__restrict__ ims, const double *__restrict__ sums, float *__restrict__ x, long long HW, float rgb_max) {
    vv_pdl_wait();
    const int bc = blockIdx.y, b = bc / 3, c = bc - 3 * b;
    const float mean = (float)(sums[bc] / (double)(2 * HW));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * HW; i += (long long)gridDim.x * blockDim.x) {
        const int f = i >= HW;
        const long long pix = i - f * HW;
        x[((long long)b * 6 + f * 3 + c) * HW + pix] = (ims[(long long)bc * 2 * HW + i] - mean) / rgb_max;
    }
}

// out[b, c, Y, X] = mul * upsample4(in)[b, c, Y, X];  mode 0: bilinear, align_corners = False (F.interpolate's default, which is
// what nn.Upsample(scale_factor=4, mode='bilinear') resolves to in PyTorch >= 0.4); mode 1: nearest
__global__ void k_fn_upsample4(const float *__restrict__ in, long long in_bs, int C, int h, int w, float *__restrict__ out, long long out_bs,
                               int mode, float mul, int B) {
    vv_pdl_wait();
    const int H = 4 * h, W = 4 * w;
    const long long total = (long long)B * C * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(i % W);
        long long r = i / W;
        const int Y = (int)(r % H); r /= H;
        const int c = (int)(r % C), b = (int)(r / C);
        const float *src = in + (long long)b * in_bs + (long long)c * h * w;
        float v;
        if (mode == 1) {
            v = src[(Y >> 2) * w + (X >> 2)];
        } else {
            float sy = 0.25f * ((float)Y + 0.5f) - 0.5f, sx = 0.25f * ((float)X + 0.5f) - 0.5f;
            sy = sy < 0.f ? 0.f : sy; sx = sx < 0.f ? 0.f : sx;
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
            const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
            v = hy * (hx * src[y0 * w + x0] + lx * src[y0 * w + x1]) + ly * (hx * src[y1 * w + x0] + lx * src[y1 * w + x1]);
        }
        out[(long long)b * out_bs + ((long long)c * H + Y) * W + X] = mul * v;
    }
}

// out slice = act(mul * in slice); slope < 0: no activation
__global__ void k_fn_scale_copy(const float *__restrict__ in, long long in_bs, float *__restrict__ out, long long out_bs, long long per,
                                float mul, float slope, int B) {
    vv_pdl_wait();
    const long long total = (long long)B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        const long long r = i - (long long)b * per;
        float v = mul * in[(long long)b * in_bs + r];
        if (slope >= 0.f) v = v > 0.f ? v : slope * v;
        out[(long long)b * out_bs + r] = v;
    }
}

// the launch description of a Conv2d / of the four fused phases of a ConvTranspose2d(4, 2, 1), minus pointers
FnConv describe_conv(int c_in, int in_h, int in_w, int c_out, int ksize, int stride, int batch) {
    const int pad = (ksize - 1) / 2;
    FnConv p;
    memset(&p, 0, sizeof(p));
    p.Cin = c_in; p.IH = in_h; p.IW = in_w; p.Co = c_out;
    p.OH = (in_h + 2 * pad - ksize) / stride + 1; p.OW = (in_w + 2 * pad - ksize) / stride + 1;
    p.GH = p.OH; p.GW = p.OW; p.o_mul = 1; p.o_off_y = 0; p.o_off_x = 0; p.i_mul = stride;
    p.ntaps = ksize * ksize; p.kw = ksize; p.ty0 = -pad; p.dty = 1; p.tx0 = -pad; p.dtx = 1;
    p.B = batch; p.phases = 1;
    return p;
}
FnConv describe_deconv(int c_in, int in_h, int in_w, int c_out, int batch) {
    FnConv p;
    memset(&p, 0, sizeof(p));
    p.Cin = c_in; p.IH = in_h; p.IW = in_w; p.Co = c_out;
    p.OH = 2 * in_h; p.OW = 2 * in_w;
    p.GH = in_h; p.GW = in_w; p.o_mul = 2; p.o_off_y = 0; p.o_off_x = 0; p.i_mul = 1;
    // output row 2y + py receives input row y + dy through kernel row ky (oy = 2 iy - 1 + ky):
    //   py = 0: (ky, dy) = (1, 0), (3, -1);   py = 1: (ky, dy) = (0, +1), (2, 0)      -- dy = py - tap, likewise in x.
    // All four phases run in one launch (FnConv::phases): the kernel adds (py, px) to the tap origin and to the output pixel.
    p.ntaps = 4; p.kw = 2; p.ty0 = 0; p.dty = -1; p.tx0 = 0; p.dtx = -1; p.phases = 4;
    p.B = batch;
    return p;
}

}  // namespace

extern "C" int vecvad_fn_conv2d(const float *in, int64_t in_batch_stride, int c_in, int in_h, int in_w, const float *w, const float *bias,
                                float *out, int64_t out_batch_stride, int c_out, int ksize, int stride, int leaky, int batch,
                                float *scratch, int64_t scratch_floats, vecvad_stream stream) {
    VV_REQUIRE(in && w && out, "fn_conv2d: null pointer");
    VV_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "fn_conv2d: kernel size %d not in {1,3,5,7}", ksize);
    VV_REQUIRE(stride == 1 || stride == 2, "fn_conv2d: stride %d not in {1,2}", stride);
    VV_REQUIRE(batch >= 1 && c_in >= 1 && c_out >= 1 && in_h >= 1 && in_w >= 1, "fn_conv2d: bad shape");
    FnConv p = describe_conv(c_in, in_h, in_w, c_out, ksize, stride, batch);
    p.in = in; p.in_bs = in_batch_stride; p.w = w; p.bias = bias; p.out = out; p.out_bs = out_batch_stride; p.act = leaky;
    return launch_conv(p, scratch, scratch_floats, (cudaStream_t)stream);
}

// w_phases: [4][c_out][c_in * 4] -- phase (py, px) = (ph >> 1, ph & 1), tap t = a * 2 + b with the (ky, kx) pairs of
// vecvad_fn_deconv_taps(); prepared once from the ConvTranspose2d weight [c_in][c_out][4][4] (vec_vad_b200/flownet2.py)
extern "C" int vecvad_fn_deconv4x4s2(const float *in, int64_t in_batch_stride, int c_in, int in_h, int in_w, const float *w_phases,
                                     const float *bias, float *out, int64_t out_batch_stride, int c_out, int leaky, int batch,
                                     float *scratch, int64_t scratch_floats, vecvad_stream stream) {
    VV_REQUIRE(in && w_phases && out, "fn_deconv4x4s2: null pointer");
    VV_REQUIRE(batch >= 1 && c_in >= 1 && c_out >= 1 && in_h >= 1 && in_w >= 1, "fn_deconv4x4s2: bad shape");
    FnConv p = describe_deconv(c_in, in_h, in_w, c_out, batch);
    p.in = in; p.in_bs = in_batch_stride; p.w = w_phases; p.bias = bias; p.out = out; p.out_bs = out_batch_stride; p.act = leaky;
    return launch_conv(p, scratch, scratch_floats, (cudaStream_t)stream);
}

extern "C" int vecvad_fn_conv_plan(int c_in, int in_h, int in_w, int c_out, int ksize, int stride, int transposed, int batch,
                                   int64_t scratch_floats, int *tile_pixels, int *tile_channels, int *ksplit) {
    VV_REQUIRE(tile_pixels && tile_channels && ksplit, "fn_conv_plan: null pointer");
    VV_REQUIRE(batch >= 1 && c_in >= 1 && c_out >= 1 && in_h >= 1 && in_w >= 1, "fn_conv_plan: bad shape");
    if (!transposed) {
        VV_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "fn_conv_plan: kernel size %d not in {1,3,5,7}", ksize);
        VV_REQUIRE(stride == 1 || stride == 2, "fn_conv_plan: stride %d not in {1,2}", stride);
    }
    const FnConv p = transposed ? describe_deconv(c_in, in_h, in_w, c_out, batch) : describe_conv(c_in, in_h, in_w, c_out, ksize, stride, batch);
    const FnPlan plan = plan_conv(p, scratch_floats > 0, scratch_floats);
    *tile_pixels = FN_TILES[plan.tile].tm; *tile_channels = FN_TILES[plan.tile].tn; *ksplit = plan.ksplit;
    return 0;
}

// the kernel rows / columns that feed output parity 0 and 1, in tap order (see above): parity 0 -> k = 1, 3; parity 1 -> k = 0, 2
extern "C" int vecvad_fn_deconv_taps(int *k_of_parity_tap) {
    VV_REQUIRE(k_of_parity_tap, "fn_deconv_taps: null pointer");
    const int k[4] = {1, 3, 0, 2};
    for (int i = 0; i < 4; i++) k_of_parity_tap[i] = k[i];
    return 0;
}

extern "C" int vecvad_fn_normalize_pair(const float *ims, float *x, double *scratch, int batch, int height, int width, float rgb_max,
                                        vecvad_stream stream) {
    VV_REQUIRE(ims && x && scratch && batch >= 1 && height >= 1 && width >= 1 && rgb_max > 0.f, "fn_normalize_pair: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const long long HW = (long long)height * width;
    VV_CK(cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * batch, st));
    const int gx = (int)((2 * HW + 256 * 8 - 1) / (256 * 8));
    VV_CK(vv_launch(k_fn_mean, dim3(gx < 1 ? 1 : gx, 3 * batch), dim3(256), 0, st, ims, scratch, 2 * HW));
    VV_CKL();
    VV_CK(vv_launch(k_fn_normalize, dim3(gx < 1 ? 1 : gx, 3 * batch), dim3(256), 0, st, ims, (const double *)scratch, x, HW, rgb_max));
    VV_CKL();
    return 0;
}

extern "C" int vecvad_fn_upsample4(const float *in, int64_t in_batch_stride, int channels, int h, int w, float *out, int64_t out_batch_stride,
                                   int mode, float mul, int batch, vecvad_stream stream) {
    VV_REQUIRE(in && out && channels >= 1 && h >= 1 && w >= 1 && batch >= 1 && (mode == 0 || mode == 1), "fn_upsample4: bad argument");
    const long long total = (long long)batch * channels * 16 * h * w;
    const int gx = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    VV_CK(vv_launch(k_fn_upsample4, dim3(gx), dim3(256), 0, (cudaStream_t)stream, in, (long long)in_batch_stride, channels, h, w, out,
                    (long long)out_batch_stride, mode, mul, batch));
    VV_CKL();
    return 0;
}

extern "C" int vecvad_fn_scale_copy(const float *in, int64_t in_batch_stride, float *out, int64_t out_batch_stride, int64_t elems_per_image,
                                    float mul, float leaky_slope, int batch, vecvad_stream stream) {
    VV_REQUIRE(in && out && elems_per_image >= 1 && batch >= 1, "fn_scale_copy: bad argument");
    const long long total = (long long)batch * elems_per_image;
    const int gx = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    VV_CK(vv_launch(k_fn_scale_copy, dim3(gx), dim3(256), 0, (cudaStream_t)stream, in, (long long)in_batch_stride, out,
                    (long long)out_batch_stride, (long long)elems_per_image, mul, leaky_slope, batch));
    VV_CKL();
    return 0;
}
