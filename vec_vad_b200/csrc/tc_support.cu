// Which implicit-GEMM problems the tcgen05 tile kernels accept at all (alignment, channel granularity, contiguous groups).
// The per-kernel planners (igemm_flat.cu, igemm_tc3.cu, wgrad_flat.cu, wgrad_tc2.cu) refine this with their own shared-memory
// plans; anything refused here or there runs on the fp32 SIMT tiles (igemm_simt.cu).
#include "tc_common.cuh"

bool vv_igemm_tc_supported(const VvIGemm &p) {
    int bw, bh, bn;
    if (!encode_fn()) return false;
    const int al = p.ab_f16 ? 8 : 4;                     // elements per 16 bytes
    if (p.Kt % KS || p.N % 32 || p.N < 32) return false;
    if (p.lda % al || p.a_coff % al || p.ldo % 4 || p.o_coff % 4) return false;
    if (((uintptr_t)p.A) % 16 || ((uintptr_t)p.Wt) % 16 || ((uintptr_t)p.O) % 16) return false;
    if (p.a_s2d && ((p.Kt / 4) % KS)) return false;
    if (p.o_d2s && ((p.N / 4) % 32)) return false;
    if (p.o_split && (p.o_split % 32 || !p.O2 || p.ldo2 % 8 || ((uintptr_t)p.O2) % 16 || p.o_d2s)) return false;
    if (p.G > 1 && (p.a_gs != (long long)p.B * p.H * p.W * (p.a_s2d ? 4 : 1) * p.lda)) return false;   // groups must be contiguous images
    if (p.G > 1 && p.w_gs != (long long)p.taps.n * p.N * p.Kt) return false;
    return tile_geometry_n(p.H, p.W, BM, bw, bh, bn);
}

bool vv_wgrad_tc_supported(const VvWGrad &p) {
    int bw, bh, bn;
    if (!encode_fn()) return false;
    const int al = p.ab_f16 ? 8 : 4;
    if (p.Kt % KS || p.N % 32 || p.N < 32) return false;
    if (p.lda % al || p.a_coff % al || p.ldg % al || p.g_coff % al || p.a_gs % al || p.g_gs % al) return false;
    if (((uintptr_t)p.A) % 16 || ((uintptr_t)p.Gd) % 16) return false;
    if (p.g_s2d && ((p.N / 4) % KS)) return false;
    return tile_geometry_n(p.H, p.W, BM, bw, bh, bn);
}
