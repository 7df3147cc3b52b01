/* vecvad.h -- C ABI of libvecvad.so, the B200 (sm_100a) implementation of the VEC_VAD hot path.
 *
 * Plain C: raw device pointers, explicit sizes, a cudaStream_t passed as void*, int status.
 * The library never allocates device memory behind the caller's back: every buffer
 * (parameters, gradients, workspace, outputs) is owned by the caller (PyTorch in this repo,
 * see vec_vad_b200/_lib.py) and handed in as a pointer.  Every entry point returns 0 on
 * success and a negative code on failure; vecvad_last_error() gives the message (the reference
 * FFI printed and aborted instead: correlation_cuda_kernel.cu:362-368, correlation_cuda.c:87-89).
 *
 * Two groups of entry points:
 *   (1) the FlowNet2 ops, replacing the reference's cffi symbols one for one
 *         Correlation_forward_cuda / _backward_cuda   (ops/correlation/src/correlation_cuda.h:1-17)
 *         Resample2d_cuda_forward / _backward         (ops/resample2d/src/Resample2d_cuda.h:1-3)
 *         ChannelNorm_cuda_forward / _backward        (ops/channelnorm/src/ChannelNorm_cuda.h:1-3)
 *   (2) the completion-UNet set (model/unet.py:73-652) + train-step body (train.py:383-402),
 *       which in the reference is PyTorch/cuDNN called from Python: the binding surface is the
 *       nn.Module API (vec_vad_b200/unet.py); these are the native calls underneath it.
 */
#ifndef VECVAD_H_
#define VECVAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VECVAD_ABI_VERSION 8
#define VECVAD_MAX_UNETS 10   /* 5 raw + 5 flow (SelfCompleteNetFull, model/unet.py:270-408) */
#define VECVAD_N_UNITS 14     /* conv3x3+BN+ReLU units per UNet (model/unet.py:187-196) */
#define VECVAD_N_UPS 3        /* ConvTranspose2d per UNet (model/unet.py:54) */

typedef void *vecvad_stream;  /* cudaStream_t */

int vecvad_abi_version(void);
const char *vecvad_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py: gpu_launches) */
uint64_t vecvad_launch_count(void);

/* per-kernel-class device timing (CUDA events on the launch stream) for bench.py's roofline leg.
 * classes: 0 conv/dgrad tcgen05 tiles, 1 wgrad tcgen05 tiles, 2 conv/dgrad fp32 SIMT tiles, 3 wgrad fp32 SIMT tiles,
 *          4 BatchNorm apply/backward passes, 5 (unused).  begin() resets and enables; end() synchronises the device, disables,
 * and returns per class the summed milliseconds, algorithmic FLOPs and launch counts since begin(). */
#define VECVAD_PROFILE_CLASSES 6
int vecvad_profile_begin(void);
int vecvad_profile_end(double *ms, double *flops, int64_t *launches, int n_classes);

/* ------------------------------------------------------------------------------------------
 * (1) FlowNet2 ops.  All tensors are contiguous NCHW fp32 on the current device (the reference
 * asserts contiguity: functions/correlation.py:17-18, functions/resample2d.py:9-10).
 * Unlike the reference (which resizes + zero-fills tensors it is handed, correlation_cuda.c:36-42)
 * the caller allocates the output; vecvad_correlation_out_shape() gives its size
 * (shape rule: correlation_cuda.c:25-34).  No padded NHWC scratch copies (rInput1/rInput2) exist.
 * ------------------------------------------------------------------------------------------ */
int vecvad_correlation_out_shape(int in_h, int in_w, int pad_size, int kernel_size, int max_displacement,
                                 int stride1, int stride2, int *out_c, int *out_h, int *out_w);

/* scratch bytes the fast (FlowNetC-parameter, TMA-fed) forward kernel wants: two column-parity-split copies of the inputs.
 * 0 when the parameters are served by the general kernel, which needs none. */
int vecvad_correlation_workspace_bytes(int batch, int channels, int in_h, int in_w, int pad_size, int kernel_size,
                                       int max_displacement, int stride1, int stride2, int64_t *bytes);

/* replaces Correlation_forward_cuda (correlation_cuda.c:11-93; kernels correlation_cuda_kernel.cu:10-106).
 * out[n,tc,y,x] = 1/(k*k*C) * sum_{j,i,c} in1[n,c,y1+j,x1+i] * in2[n,c,y1+tj*s2+j,x1+ti*s2+i] (zero padded).
 * workspace (16-byte aligned, >= vecvad_correlation_workspace_bytes) may be NULL: the general kernel is used then. */
int vecvad_correlation_forward(const float *in1, const float *in2, float *out, int batch, int channels, int in_h, int in_w,
                               int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                               int corr_type_multiply, void *workspace, int64_t workspace_bytes, vecvad_stream stream);

/* replaces Correlation_backward_cuda (correlation_cuda.c:95-180; kernels :108-290). grad_in1/2 are fully written. */
int vecvad_correlation_backward(const float *in1, const float *in2, const float *grad_out, float *grad_in1, float *grad_in2,
                                int batch, int channels, int in_h, int in_w, int pad_size, int kernel_size,
                                int max_displacement, int stride1, int stride2, int corr_type_multiply, vecvad_stream stream);

/* replaces Resample2d_cuda_forward (Resample2d_kernel.cu:20-66,188-206): border-clamped bilinear warp.
 * img [B,C,H,W], flow [B,2,Ho,Wo] -> out [B,C,Ho,Wo] (reference allocates out with the flow's H,W:
 * functions/resample2d.py:16-19). */
int vecvad_resample2d_forward(const float *img, const float *flow, float *out, int batch, int channels, int img_h, int img_w,
                              int out_h, int out_w, int kernel_size, vecvad_stream stream);

/* replaces Resample2d_cuda_backward (Resample2d_kernel.cu:69-186,208-238). grad_img is zeroed then scatter-added. */
int vecvad_resample2d_backward(const float *img, const float *flow, const float *grad_out, float *grad_img, float *grad_flow,
                               int batch, int channels, int img_h, int img_w, int out_h, int out_w, int kernel_size,
                               vecvad_stream stream);

/* replaces ChannelNorm_cuda_forward (ChannelNorm_kernel.cu:19-51): out[b,0,y,x] = sqrt(sum_c in^2). */
int vecvad_channelnorm_forward(const float *in, float *out, int batch, int channels, int h, int w, int norm_deg,
                               vecvad_stream stream);

/* replaces ChannelNorm_cuda_backward (ChannelNorm_kernel.cu:54-81): g*x/(out+1e-9). */
int vecvad_channelnorm_backward(const float *in, const float *out, const float *grad_out, float *grad_in, int batch,
                                int channels, int h, int w, int norm_deg, vecvad_stream stream);

/* fused img0 - warp(img1, flow) and its channel norm (FlowNet2 call sites flownet2.py:79-81,93-95,108-115):
 * warped [B,C,H,W], diff [B,C,H,W] (may be NULL), norm [B,1,H,W] (may be NULL). */
int vecvad_warp_diff_norm(const float *img0, const float *img1, const float *flow, float *warped, float *diff, float *norm,
                          int batch, int channels, int h, int w, vecvad_stream stream);

/* ------------------------------------------------------------------------------------------
 * (2) Completion-UNet set.
 *
 * A "net" is G independent UNets that read the same 5-frame cube batch x[B, 3*tot_raw, S, S]
 * (NCHW fp32, channel = 3*t+c; vad_datasets.py:159-160).  UNet g sees the cube with frame
 * erase_frame[g] dropped (padding=0, model/unet.py:183) or zeroed (padding=1, :180-181) and
 * regresses either that raw frame (3 channels) or a flow frame of x_of (2 channels).
 *
 * Parameters live in ONE flat fp32 buffer owned by the caller, laid out per UNet "slot" with a
 * uniform stride; inside a slot every tensor keeps PyTorch's own layout (Conv2d [Cout,Cin,3,3],
 * ConvTranspose2d [Cin,Cout,3,3]) at the offsets given below, so nn.Parameter views alias it.
 * Gradients use a second buffer with the identical layout; BatchNorm running statistics a third.
 * ------------------------------------------------------------------------------------------ */
typedef struct vecvad_net_config {
    int n_unets;                              /* G: UNets executed per forward (<= VECVAD_MAX_UNETS)            */
    int features_root;                        /* nf (config.cfg:62); multiple of 16                              */
    int tot_raw_num;                          /* frames per cube (5)                                             */
    int patch;                                /* S: patch size (32); multiple of 8                               */
    int padding;                              /* 0: erased frame dropped, 1: erased frame zeroed                 */
    int param_slot[VECVAD_MAX_UNETS];         /* which slot of the flat buffers UNet g uses                      */
    int erase_frame[VECVAD_MAX_UNETS];        /* frame removed from UNet g's input                               */
    int out_channels[VECVAD_MAX_UNETS];       /* 3 (raw) or 2 (flow)                                             */
    int target_is_flow[VECVAD_MAX_UNETS];     /* 0: target = x[:, 3*target_index ...]; 1: x_of[:, 2*target_index]*/
    int target_index[VECVAD_MAX_UNETS];
    int out_slot[VECVAD_MAX_UNETS];           /* position inside raw_out (3 ch each) or of_out (2 ch each)       */
    int64_t slot_param_stride;                /* floats between consecutive slots in params / grads              */
    int64_t slot_stat_stride;                 /* floats between consecutive slots in the running-stat buffer     */
    /* offsets (floats) inside one slot */
    int64_t conv_w[VECVAD_N_UNITS], conv_b[VECVAD_N_UNITS], bn_w[VECVAD_N_UNITS], bn_b[VECVAD_N_UNITS];
    int64_t up_w[VECVAD_N_UPS], up_b[VECVAD_N_UPS];
    int64_t out_w, out_b;
    int64_t run_mean[VECVAD_N_UNITS], run_var[VECVAD_N_UNITS];
    int use_tensor_cores;                     /* 0: fp32 SIMT tiles; 1: tcgen05 kind::tf32 implicit-GEMM tiles (fp32 tensors in HBM, operands
                                               * rounded to tf32 on their way into shared memory); 2: tcgen05 kind::f16 tiles: post-BN
                                               * activations, dZ gradients and re-laid-out weights are stored as fp16 in HBM (the same 10-bit
                                               * mantissa, rounded once where they are produced), accumulation, BatchNorm statistics, raw conv
                                               * outputs, losses and the optimiser stay fp32 */
    /* A UNet set may be split over several nets (one per stream) that share the flat buffers and the output tensors:
     * the loss means and the external gradient tensors then span ALL raw / flow outputs, not only this net's.
     * 0 = this net's own count. */
    int n_raw_total, n_of_total;
} vecvad_net_config;

typedef struct vecvad_net vecvad_net;

int vecvad_net_create(const vecvad_net_config *cfg, vecvad_net **out);
void vecvad_net_destroy(vecvad_net *net);

/* bytes of device workspace needed for a batch of `batch` cubes (activations saved for backward,
 * gradient scratch, re-laid-out weights). */
int vecvad_net_workspace_bytes(const vecvad_net *net, int batch, int64_t *bytes);

/* bind caller-owned device buffers. running_stats may be NULL only if the net is never run in
 * training mode and never in eval mode (i.e. never). workspace must be 256-byte aligned. */
int vecvad_net_bind(vecvad_net *net, float *params, float *grads, float *running_stats, void *workspace,
                    int64_t workspace_bytes, int max_batch);

/* forward.  x [B,3*tot_raw,S,S], x_of [B,2*T_of,S,S] (may be NULL when no UNet targets flow).
 * raw_out [B,3*n_raw_out,S,S] / of_out [B,2*n_of_out,S,S]: UNet g writes its channels at out_slot[g].
 * training=1: batch statistics, running stats updated (momentum 0.1, unbiased var; nn.BatchNorm2d defaults),
 *             activations kept for vecvad_net_backward.   training=0: running statistics.
 * If sse (device, [G][B] floats) is non-NULL the per-cube sum of squared error against the targets is written
 * (train.py:414-427 scoring), and -- when training -- d(loss)/d(out) is staged for vecvad_net_backward with
 *   loss = lambda_raw * mean((raw_tgt-raw_out)^2) + lambda_of * mean((of_tgt-of_out)^2)      (train.py:385-392). */
int vecvad_net_forward(vecvad_net *net, const float *x, const float *x_of, int x_of_channels, int batch, int training,
                       float *raw_out, int raw_out_channels, float *of_out, int of_out_channels, float *sse,
                       float lambda_raw, float lambda_of, vecvad_stream stream);

/* backward of the last training forward.  grad_raw_out/grad_of_out: NCHW gradients wrt the outputs, or both
 * NULL to use the fused MSE gradient staged by vecvad_net_forward(..., sse != NULL).
 * Writes (overwrites) every gradient of the executed UNets into the bound `grads` buffer. */
int vecvad_net_backward(vecvad_net *net, const float *grad_raw_out, const float *grad_of_out, vecvad_stream stream);

/* Gradient phases (overlapping the data-parallel gradient exchange with the backward; replaces the reduce-add onto GPU 0 that
 * nn.DataParallel performs after the whole backward, train.py:375,399).  The parameter gradients of every slot become final
 * in three contiguous ranges of the slot, in this order: phase 0 = the decoder (conv units 8..13, transposed convs, output conv),
 * phase 1 = the deepest encoder block (units 6, 7), phase 2 = the rest.
 * vecvad_net_grad_phase_ranges: [begin[p], end[p]) in floats relative to the start of a slot (same for every slot).
 * vecvad_net_grad_phase_wait:   makes `stream` wait (cudaStreamWaitEvent) until phase p of the LAST vecvad_net_backward call is
 *                               complete; call it after vecvad_net_backward returned (the backward itself is asynchronous). */
int vecvad_net_grad_phase_ranges(const vecvad_net *net, int64_t *begin, int64_t *end);
/* defer != 0: vecvad_net_backward returns WITHOUT making its stream wait for the side stream's last weight gradients; the caller must
 * call vecvad_net_grad_phase_wait(net, 2, stream) before it reads phase-2 gradients or runs the next forward.  In between it can work on
 * phases 0 and 1 (vecvad_net_grad_phase_wait for each, then e.g. vecvad_adam_step_ranges) while those last weight gradients finish. */
int vecvad_net_defer_join(vecvad_net *net, int defer);
int vecvad_net_grad_phase_wait(vecvad_net *net, int phase, vecvad_stream stream);

/* fp16-operand mode only: the power-of-two loss scale the backward applies to the dZ operands it stores as fp16 (and removes again
 * from every parameter gradient it writes).  scale <= 0 (default): derived per step from the MSE-gradient coefficients of the last
 * forward (2*lambda / elements), which suits gradients of the size the MSE losses of train.py:385-392 produce; pass an explicit
 * power of two when external output gradients of a very different magnitude are fed to vecvad_net_backward. */
int vecvad_net_set_loss_scale(vecvad_net *net, float scale);

/* losses[0] = mean raw MSE, losses[1] = mean flow MSE (0 if no flow UNet) from the sse buffer of the last forward. */
int vecvad_net_losses(vecvad_net *net, const float *sse, int batch, float *losses, vecvad_stream stream);

/* torch.optim.Adam semantics (train.py:376: lr 1e-3, betas (0.9,0.999), eps 1e-7, weight_decay 0) on flat buffers.
 * `step` is the 1-based step count used for bias correction.  grad_scale multiplies g first (1/world_size after a
 * summing all-reduce). */
int vecvad_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, vecvad_stream stream);
/* The same update restricted to [begin, end) (floats) of each of n_slots slots of the flat buffers, slot_stride floats apart: with the
 * ranges of vecvad_net_grad_phase_ranges the optimiser can start on the phases that are final while the last weight gradients still
 * run (see vecvad_net_defer_join). */
int vecvad_adam_step_ranges(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int n_slots, int64_t slot_stride,
                            int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                            float grad_scale, vecvad_stream stream);

/* debug / test facility: copy an internal workspace buffer (device to device) into dst as fp32.  kind: 0 X0, 1 Z[u], 2 A[u] (even u),
 * 3 CAT[k], 4 PL[k], 5 X4, 6 UU[k], 7 dCAT[k], 8 GA, 9 GB, 10 DOUT, 11 Wf[u], 12 dWf[u], 13 tWf[k], 14 tdW[k], 15 dUP[k] and
 * 16 dZ ring buffer k (fp16 mode only).  Buffers are grouped NHWC [G][B*H*W][C] of the last forward; fp16 buffers of the fp16 mode
 * are converted; *n_floats receives the element count copied. */
int vecvad_net_debug_read(vecvad_net *net, int kind, int index, float *dst, int64_t max_floats, int64_t *n_floats,
                          vecvad_stream stream);

/* ---- single ops on NHWC tensors, exported for unit tests and for profiling one kernel at a time ---- */

/* use_tc of every single op: 0 = fp32 SIMT tiles; low four bits 1 = the tcgen05 tile the net engine picks for the shape,
 * 3 = flattened-sequence tiles (what the engine uses at >= 32 pixels per row), 4 (forward / dgrad) or 2 (wgrad) = pair / tap-reuse
 * tiles (what it uses below that); + 16 = fp16 operands (kind::f16, fp32 accumulation): inputs are converted to fp16 inside the
 * call into stream-ordered temporaries, outputs stay fp32. */

/* 3x3 pad-1 convolution as implicit GEMM.  in [B,H,W,cin] (row stride ld_in), w [cout,cin,3,3] PyTorch layout,
 * out [B,H,W,cout] raw (pre-BN) values; stats[2*cout] (double) receives per-channel sum and sum of squares
 * (may be NULL).  scratch: >= 9*cout*cin floats. */
int vecvad_conv3x3_forward(const float *in, int ld_in, const float *w, const float *bias, float *out, double *stats,
                           float *scratch, int batch, int h, int wd, int cin, int cout, int use_tc, vecvad_stream stream);

/* weight gradient of the same convolution: dw [cout,cin,3,3] (PyTorch layout, overwritten) from in [B,H,W,cin] and
 * grad_out [B,H,W,cout] (NHWC, dense).  scratch: >= 9*cout*cin floats. */
int vecvad_conv3x3_wgrad(const float *in, int ld_in, const float *grad_out, float *dw, float *scratch, int batch, int h, int wd,
                         int cin, int cout, int use_tc, vecvad_stream stream);

/* input gradient of the same convolution (what the engine's backward launches for every unit but the first): grad_out
 * [B,H,W,cout] dense NHWC, w [cout,cin,3,3] -> grad_in [B,H,W,cin].  scratch: >= 18*cout*cin floats.  use_tc: 0 fp32 SIMT tiles,
 * nonzero: the tcgen05 tile the engine picks for this shape (flattened-sequence tiles at >= 32 pixels per row, pair tiles below). */
int vecvad_conv3x3_dgrad(const float *grad_out, const float *w, float *grad_in, float *scratch, int batch, int h, int wd, int cin,
                         int cout, int use_tc, vecvad_stream stream);

/* ConvTranspose2d(ci -> co, kernel 3, stride 2, padding 1, output_padding 1) (model/unet.py:54) exactly as the engine runs it:
 * the four sub-pixel phases as one 2x2-tap contraction over 4*co columns, pixel-shuffled by the epilogue into
 * out [B,2H,2W,ld_out] at channel offset out_coff (the second half of the concat buffer, model/unet.py:59).
 * in [B,H,W,ci] dense NHWC, w [ci,co,3,3] PyTorch layout, bias [co].  scratch: >= 32*co*ci + co floats (wgrad: 48*co*ci + co).
 * _dgrad: grad_out [B,2H,2W,ld] (channels coff .. coff+co) -> grad_in [B,H,W,ci];  _wgrad: -> dw [ci,co,3,3] (overwritten). */
int vecvad_convt3x3s2_forward(const float *in, const float *w, const float *bias, float *out, int ld_out, int out_coff,
                              float *scratch, int batch, int h, int wd, int ci, int co, int use_tc, vecvad_stream stream);
int vecvad_convt3x3s2_dgrad(const float *grad_out, int ld, int coff, const float *w, float *grad_in, float *scratch, int batch,
                            int h, int wd, int ci, int co, int use_tc, vecvad_stream stream);
int vecvad_convt3x3s2_wgrad(const float *in, const float *grad_out, int ld, int coff, float *dw, float *scratch, int batch, int h,
                            int wd, int ci, int co, int use_tc, vecvad_stream stream);

/* ---- FlowNet2 inference graph (FlowNet2_src/models/flownet2.py:65-149, components/FlowNet{C,S,SD,Fusion}.py, misc.py:6-45): the layers
 * that are not the three native ops.  NCHW fp32; every tensor argument is a CHANNEL SLICE of a possibly larger buffer (pointer to the
 * slice's first channel + the batch stride, in elements, of the buffer it lives in), so the reference's torch.cat calls never run.
 * fn_conv2d        nn.Conv2d(c_in, c_out, ksize, stride, padding=(ksize-1)/2, bias) [+ LeakyReLU(0.1) when leaky]   (misc.py:6-27,41-45)
 *                  w [c_out][c_in][k][k] (PyTorch layout), bias nullable.
 * fn_deconv4x4s2   nn.ConvTranspose2d(c_in, c_out, 4, 2, 1, bias) [+ LeakyReLU(0.1)]                              (misc.py:30-38)
 *                  w_phases [4][c_out][c_in*4]: the weight [c_in][c_out][4][4] re-laid out once per output parity phase
 *                  (py, px) = (ph >> 1, ph & 1), tap (a, b) -> kernel element (k[py*2+a], k[px*2+b]) with k from fn_deconv_taps.
 * fn_normalize_pair ims [B,3,2,H,W] -> x [B,6,H,W] = (ims - mean over both frames per colour) / rgb_max        (flownet2.py:66-72)
 *                  scratch: 3*B doubles.
 * fn_upsample4     nn.Upsample(scale_factor=4, mode = 0 'bilinear' (align_corners False) | 1 'nearest'), times mul.
 * fn_scale_copy    out slice = mul * in slice, then LeakyReLU(leaky_slope) if leaky_slope >= 0.
 * scratch (fn_conv2d / fn_deconv4x4s2): caller-owned device floats (may be NULL): launches whose pixel grid is too small to fill the
 *                  GPU split the contraction over more CTAs and pass their partial sums through it (used size <= scratch_floats).
 * fn_conv_plan     host only, no launch: the CTA tile (pixels x channels) and the split of the contraction that fn_conv2d
 *                  (transposed = 0) or fn_deconv4x4s2 (transposed = 1; ksize / stride ignored) uses for this shape with a scratch
 *                  buffer of scratch_floats floats (0: none).  A function of the arguments alone -- so is the summation order. */
int vecvad_fn_conv2d(const float *in, int64_t in_batch_stride, int c_in, int in_h, int in_w, const float *w, const float *bias, float *out,
                     int64_t out_batch_stride, int c_out, int ksize, int stride, int leaky, int batch, float *scratch,
                     int64_t scratch_floats, vecvad_stream stream);
int vecvad_fn_deconv4x4s2(const float *in, int64_t in_batch_stride, int c_in, int in_h, int in_w, const float *w_phases, const float *bias,
                          float *out, int64_t out_batch_stride, int c_out, int leaky, int batch, float *scratch, int64_t scratch_floats,
                          vecvad_stream stream);
int vecvad_fn_deconv_taps(int *k_of_parity_tap);
int vecvad_fn_conv_plan(int c_in, int in_h, int in_w, int c_out, int ksize, int stride, int transposed, int batch, int64_t scratch_floats,
                        int *tile_pixels, int *tile_channels, int *ksplit);
int vecvad_fn_normalize_pair(const float *ims, float *x, double *scratch, int batch, int height, int width, float rgb_max, vecvad_stream stream);
int vecvad_fn_upsample4(const float *in, int64_t in_batch_stride, int channels, int h, int w, float *out, int64_t out_batch_stride, int mode,
                        float mul, int batch, vecvad_stream stream);
int vecvad_fn_scale_copy(const float *in, int64_t in_batch_stride, float *out, int64_t out_batch_stride, int64_t elems_per_image, float mul,
                         float leaky_slope, int batch, vecvad_stream stream);

/* get_foreground on the device (vad_datasets.py:70-93): crop every box out of each of n_frames frames and resize the crop to
 * patch x patch with the arithmetic of cv2.resize(.., INTER_LINEAR), bit-exact for uint8 and float32 frames alike (copy when the
 * crop already has the size, 2x2 box average when it is exactly twice as large, fixed-point / unfused float bilinear otherwise).
 * frames: device, uint8 (is_f32 = 0) or float32 (is_f32 = 1), addressed as [t][c][y][x] through the four ELEMENT strides, so cv2's
 *         [T,H,W,C] frames (stride_c = 1) and the reference's [T,C,H,W] stacks both pass unchanged.
 * boxes : device int32 [n_boxes][4] = x_min, y_min, x_max, y_max, the ceil'ed bbox edges exactly as the reference computes them on the
 *         host (np.ceil -> int); every box must satisfy 0 <= min < max <= extent (checked by the host wrapper).
 * out   : device [n_boxes][n_frames][channels][patch][patch], same element type as frames; patch <= 32. */
int vecvad_crop_resize(const void *frames, int is_f32, int n_frames, int channels, int height, int width, int64_t stride_t,
                       int64_t stride_c, int64_t stride_h, int64_t stride_w, const int32_t *boxes, int n_boxes, int patch, void *out,
                       vecvad_stream stream);

/* cube staging: uint8 cubes [N,T,S,S,3] (+ float flow [N,T_of,S,S,2]) -> x [N,3T,S,S] float /255, x_of [N,2*T_of,S,S]
 * == cube_to_train_dataset + ToTensor + collate (vad_datasets.py:130-168). */
int vecvad_cubes_to_tensors(const uint8_t *raw, const float *flow, float *x, float *x_of, int n, int t_raw, int t_of, int patch,
                            vecvad_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VECVAD_H_ */
