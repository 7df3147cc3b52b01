// Parameter blocks + launcher prototypes of the bandwidth-bound UNet kernels (unet_kernels.cu).
#pragma once
#include <math.h>

#include "common.h"

struct VvIntG {
    int v[VECVAD_MAX_UNETS];
};

struct VvBnApply {
    const float *Z;      long long z_gs;                  // [G][M][C] raw conv output
    float *Y;            long long y_gs;  int ldy, y_coff; // destination view (may live inside a concat buffer)
    float *P;            long long p_gs;                  // pooled destination [G][M/4][C] (pool != 0)
    int pool;
    int M, H, W, C;
    int training;
    const double *stats; long long stats_gs;              // [G][2][C] sums (training)
    const float *vec;    long long vec_gs;                // [G][3][C] bias, gamma, beta
    float *running;      VvIntG slot;  long long slot_stat_stride, rm_off, rv_off;
    float *save;         long long save_gs;               // [G][4][C] scale, shift, mean, invstd (training)
};

struct VvBnBwd {
    const float *Z;      long long z_gs;
    const float *dY;     long long dy_gs; int ldy, dy_coff;
    float *dZ;           long long dz_gs;                 // dense [G][M][C] (may alias dY when dY is dense)
    int M, C;
    const float *save;   long long save_gs;
    double *sums;        long long sums_gs;               // [G][2][C], pre-zeroed
    float *grads;        VvIntG slot;  long long slot_param_stride, gamma_off, beta_off;
    // fused 1x1 output conv backward (last unit of the UNet, C == features_root): dY[m][c] = sum_j dout[m][j] * w_out[j][c] is formed on
    // the fly from the staged loss gradient instead of being read, and the 1x1 conv's own weight / bias gradients are reduced here
    const float *dout;   const float *params;  long long ow_off, ob_off;  VvIntG out_channels;      // dout [G][M][4]; NULL = plain mode
};

struct VvOutFwd {
    const float *U;      long long u_gs;                  // [G][B*S*S][F]
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
    const float *U;      long long u_gs;
    float *dU;           long long du_gs;
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
    float *Wf, *Wd, *vec;          // Wd may be NULL
    const float *dWf;              // scatter direction
    long long wf_gs, wd_gs, vec_gs;
    int blk0;                      // first block of this unit
};
struct VvPrepAll {
    VvPrepUnit u[VECVAD_N_UNITS];
    int n, total_blocks;
};
int vv_prep_conv_w_all(const float *params, const VvIntG &slot, long long slot_stride, VvPrepAll &all, int G, cudaStream_t st);
int vv_scatter_conv_wgrad_all(float *grads, const VvIntG &slot, long long slot_stride, VvPrepAll &all, int G, cudaStream_t st);
int vv_prep_input(const float *x, float *X0, int G, int B, int T, int S, int cinp, int padding, const VvIntG &erase, cudaStream_t st);
int vv_prep_conv_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, long long g_off,
                   long long beta_off, int N, int C, int Cp, float *Wf, long long wf_gs, float *Wd, long long wd_gs, float *vec,
                   long long vec_gs, int G, cudaStream_t st);
int vv_prep_ct_w(const float *params, const VvIntG &slot, long long slot_stride, long long w_off, long long b_off, int Ci, int Co,
                 float *Wbf, long long wf_gs, float *Wbd, long long wd_gs, float *vec, long long vec_gs, int G, cudaStream_t st);
int vv_bn_apply(const VvBnApply &p, int G, cudaStream_t st);
int vv_bn_bwd(const VvBnBwd &p, int G, cudaStream_t st);
int vv_maxpool_bwd(const float *Y, long long y_gs, int ldy, int y_coff, const float *dP, long long dp_gs, float *dY, long long dy_gs,
                   int lddy, int dy_coff, int G, int B, int H, int W, int C, cudaStream_t st);
int vv_colsum(const float *D, long long d_gs, int ld, int coff, int M, int C, float *grads, const VvIntG &slot, long long slot_stride,
              long long off, int G, cudaStream_t st);
int vv_outconv_fwd(const VvOutFwd &p, int G, cudaStream_t st);
int vv_outconv_bwd(const VvOutBwd &p, int G, cudaStream_t st);
int vv_scatter_conv_wgrad(const float *dWf, long long gs, int N, int C, int Cp, float *grads, const VvIntG &slot, long long slot_stride,
                          long long w_off, int G, cudaStream_t st);
int vv_scatter_ct_wgrad(const float *dWb, long long gs, int Ci, int Co, float *grads, const VvIntG &slot, long long slot_stride,
                        long long w_off, int G, cudaStream_t st);
int vv_losses(const float *sse, int G, int B, const VvIntG &is_flow, float inv_raw, float inv_of, float *out, cudaStream_t st);

// fp32 [rows][ld] (first `cols` of each row) -> dense fp16 [rows][cols], round to nearest (operand staging of the fp16 tile experiments)
int vv_f32_to_f16(const float *src, int ld, int cols, long long rows, void *dst, cudaStream_t st);

