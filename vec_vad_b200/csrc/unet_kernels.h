// Parameter blocks + launcher prototypes of the bandwidth-bound UNet kernels (unet_kernels.cu).
#pragma once
#include <math.h>

#include "common.h"

struct VvIntG {
    int v[VECVAD_MAX_UNETS];
};

struct VvBnApply {
    const void *Z;       long long z_gs;                  // [G][M][C] raw conv output: fp32, or fp16 when y_f16 (the fp16 mode stores both as fp16)
    void *Y;             long long y_gs;  int ldy, y_coff; // destination view (may live inside a concat buffer); fp32, or fp16 when y_f16
    void *P;             long long p_gs;                  // pooled destination [G][M/4][C] (pool != 0), same element type as Y
    int pool, y_f16;
    int rev;                                              // walk the rows from the end (serpentine order for L2 reuse)
    int M, H, W, C;
    int training;
    const double *stats; long long stats_gs;              // [G][2][C] sums (training)
    const float *vec;    long long vec_gs;                // [G][3][C] bias, gamma, beta
    float *running;      VvIntG slot;  long long slot_stat_stride, rm_off, rv_off;
    float *save;         long long save_gs;               // [G][4][C] scale, shift, mean, invstd (training)
};

struct VvBnBwd {
    const void *Z;       long long z_gs;  int z_f16;      // raw conv output, fp32 or fp16
    const void *dY;      long long dy_gs; int ldy, dy_coff; // same element type as Z (the fp16 mode stores output gradients as fp16 too)
    void *dZ;            long long dz_gs;                 // dense [G][M][C]: fp32 (may alias dY when dY is dense) or fp16 (dz_f16; never aliases)
    int dz_f16;
    int rev_reduce, rev_apply;                            // row order of the two passes (serpentine order for L2 reuse)
    float store_scale;                                    // dZ is multiplied by this when stored (the fp16 path's loss scale, applied once, at the last unit)
    float grad_unscale;                                   // d gamma / d beta are multiplied by this (1 / loss scale where dY already carries it)
    int M, C;
    const float *save;   long long save_gs;
    double *sums;        long long sums_gs;               // [G][2][C], pre-zeroed
    float *grads;        VvIntG slot;  long long slot_param_stride, gamma_off, beta_off;
    // fused 1x1 output conv backward (last unit of the UNet, C == features_root): dY[m][c] = sum_j dout[m][j] * w_out[j][c] is formed on
    // the fly from the staged loss gradient instead of being read, and the 1x1 conv's own weight / bias gradients are reduced here
    const float *dout;   const float *params;  long long ow_off, ob_off;  VvIntG out_channels;      // dout [G][M][4]; NULL = plain mode
};

struct VvOutFwd {
    const void *U;       long long u_gs;  int u_f16;      // [G][B*S*S][F], fp32 or fp16
    const float *params; VvIntG slot;  long long slot_param_stride, w_off, b_off;
    VvIntG out_channels, target_is_flow, target_index, out_slot;
    int B, S, F;
    float *raw_out;      int raw_out_channels;            // NCHW, nullable
    float *of_out;       int of_out_channels;
    const float *x;      int x_channels;                  // targets (NCHW), used when sse != NULL
    const float *x_of;   int x_of_channels;
    float *sse;                                           // [G][B], nullable
    float *dout;                                          // [G][B*S*S][4], nullable
    float coef_raw, coef_of;                              // d loss / d out = coef * (out - tgt)
};

struct VvOutBwd {
    const void *U;       long long u_gs;  int u_f16;
    void *dU;            long long du_gs;                 // same element type as U
    float du_scale;                                       // dU is multiplied by this (the fp16 mode's loss scale)
    const float *params; float *grads; VvIntG slot; long long slot_param_stride, w_off, b_off;
    VvIntG out_channels, target_is_flow, out_slot;
    int M, S, F;
    const float *dout;                                    // internal staged gradient [G][M][4] (used when the ext pointers are NULL)
    const float *grad_raw_out; int raw_out_channels;      // external NCHW gradients
    const float *grad_of_out;  int of_out_channels;
};

// all conv units of a net in one launch (weight re-layout forward, gradient re-layout backward)
struct VvPrepUnit {
    long long w_off, b_off, g_off, beta_off;
    int N, C, Cp;
    void *Wf, *Wd;                 // fp32 or fp16 (VvPrepAll::w_f16); Wd may be NULL
    float *vec;
    const float *dWf;              // scatter direction
    long long wf_gs, wd_gs, vec_gs;
    int blk0;                      // first block of this unit
};
struct VvPrepAll {
    VvPrepUnit u[VECVAD_N_UNITS];
    int n, total_blocks;
    int w_f16;                     // re-laid-out weights are written as fp16
    float scale;                   // scatter direction: gradients are multiplied by this (1 / loss scale)
};
int vv_prep_conv_w_all(const float *params, const VvIntG &slot, long long slot_stride, VvPrepAll &all, int G, cudaStream_t st);
int vv_scatter_conv_wgrad_all(float *grads, const VvIntG &slot, long long slot_stride, VvPrepAll &all, int G, cudaStream_t st);
int vv_prep_input(const float *x, void *X0, int x0_f16, int G, int B, int T, int S, int cinp, int padding, const VvIntG &erase, cudaStream_t st);
int vv_prep_conv_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, long long g_off,
                   long long beta_off, int N, int C, int Cp, void *Wf, long long wf_gs, void *Wd, long long wd_gs, int w_f16, float *vec,
                   long long vec_gs, int G, cudaStream_t st);
int vv_prep_ct_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, int Ci, int Co,
                 void *Wbf, long long wf_gs, void *Wbd, long long wd_gs, int w_f16, float *vec, long long vec_gs, int G, cudaStream_t st);
int vv_bn_apply(const VvBnApply &p, int G, cudaStream_t st);
int vv_bn_bwd(const VvBnBwd &p, int G, cudaStream_t st);
int vv_maxpool_bwd(const void *Y, int y_f16, long long y_gs, int ldy, int y_coff, const void *dP, long long dp_gs, void *dY, long long dy_gs,
                   int lddy, int dy_coff, int G, int B, int H, int W, int C, cudaStream_t st);
int vv_colsum(const void *D, int d_f16, long long d_gs, int ld, int coff, int M, int C, float scale, float *grads, const VvIntG &slot,
              long long slot_stride, long long off, int G, cudaStream_t st);
int vv_outconv_fwd(const VvOutFwd &p, int G, cudaStream_t st);
int vv_outconv_bwd(const VvOutBwd &p, int G, cudaStream_t st);
int vv_scatter_conv_wgrad(const float *dWf, long long gs, int N, int C, int Cp, float *grads, const VvIntG &slot, long long slot_stride,
                          long long w_off, int G, cudaStream_t st);
int vv_scatter_ct_wgrad(const float *dWb, long long gs, int Ci, int Co, float scale, float *grads, const VvIntG &slot, long long slot_stride,
                        long long w_off, int G, cudaStream_t st);
int vv_losses(const float *sse, int G, int B, const VvIntG &is_flow, float inv_raw, float inv_of, float *out, cudaStream_t st);

// fp32 [rows][ld] (first `cols` of each row) -> dense fp16 [rows][cols], round to nearest, saturating (operand staging of the single-op
// entry points in fp16 mode), and back (debug reads)
int vv_f32_to_f16(const float *src, int ld, int cols, long long rows, void *dst, cudaStream_t st);
int vv_f16_to_f32(const void *src, long long n, float *dst, cudaStream_t st);

