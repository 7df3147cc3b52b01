// TEST INFRASTRUCTURE (oracle/): a raw-pointer C harness around the REFERENCE's own FlowNet2 op kernels, compiled UNMODIFIED
// from /root/reference/FlowNet2_src/models/components/ops/*/src/*.cu (see build.sh in this directory).  It plays the part of the
// reference's THC glue, which cannot be built any more (torch.utils.ffi / THC are gone):
//   ref_correlation_*  restate correlation/src/correlation_cuda.c:11-93 (forward) and :95-180 (backward): output / scratch shapes,
//                      zero fills, then the reference launcher Correlation_{forward,backward}_cuda_kernel with contiguous strides;
//   ref_resample2d_*   restate resample2d/src/Resample2d_cuda.c:8-16 + functions/resample2d.py:8-36 (zero-filled outputs);
//   ref_channelnorm_*  restate channelnorm/src/ChannelNorm_cuda.c:8-16 + functions/channelnorm.py:8-31.
// The library built from this file (oracle/_ref/libref_ops.so) is the checker and the "recompiled reference kernels" timing bar of
// bench_flow.py.  Nothing under vec_vad_b200/ loads it.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "THC.h"
#include "correlation_cuda_kernel.h"
#include "Resample2d_kernel.h"
#include "ChannelNorm_kernel.h"

#define RCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            printf("ref_harness: %s -> %s\n", #call, cudaGetErrorString(e_));                  \
            return 0;                                                                          \
        }                                                                                      \
    } while (0)

static THCudaTensor tensor4(const float *p, long n, long c, long h, long w) {
    THCudaTensor t;
    t.data = (float *)p;
    t.size[0] = n; t.size[1] = c; t.size[2] = h; t.size[3] = w;
    t.stride[0] = c * h * w; t.stride[1] = h * w; t.stride[2] = w; t.stride[3] = 1;
    return t;
}

extern "C" {

// correlation_cuda.c:25-34
void ref_correlation_out_shape(int h, int w, int pad, int k, int md, int s1, int s2, int *oc, int *oh, int *ow) {
    const int kernel_radius = (k - 1) / 2, border_radius = kernel_radius + md;
    const int ph = h + 2 * pad, pw = w + 2 * pad;
    *oc = ((md / s2) * 2 + 1) * ((md / s2) * 2 + 1);
    *oh = (int)ceil((float)(ph - 2 * border_radius) / (float)s1);
    *ow = (int)ceil((float)(pw - 2 * border_radius) / (float)s1);
}

// scratch: two zero-padded NHWC repacks [B][H+2pad][W+2pad][C] (rInput1 / rInput2 of the reference)
long long ref_correlation_scratch_floats(int b, int c, int h, int w, int pad) { return 2LL * b * (h + 2 * pad) * (w + 2 * pad) * c; }

// returns 1 on success like the reference (correlation_cuda.c:92)
int ref_correlation_forward(const float *in1, const float *in2, float *out, float *scratch, int b, int c, int h, int w, int pad, int k,
                            int md, int s1, int s2, int corr_type_multiply, cudaStream_t stream) {
    int oc, oh, ow;
    ref_correlation_out_shape(h, w, pad, k, md, s1, s2, &oc, &oh, &ow);
    const long long rn = (long long)b * (h + 2 * pad) * (w + 2 * pad) * c;
    float *r1 = scratch, *r2 = scratch + rn;
    RCK(cudaMemsetAsync(r1, 0, 2 * rn * sizeof(float), stream));                              // THCudaTensor_fill(rInput*, 0)
    RCK(cudaMemsetAsync(out, 0, (size_t)b * oc * oh * ow * sizeof(float), stream));           // THCudaTensor_fill(output, 0)
    return Correlation_forward_cuda_kernel(out, b, oc, oh, ow, oc * oh * ow, oh * ow, ow, 1, (float *)in1, c, h, w, c * h * w, h * w, w, 1,
                                           (float *)in2, c, c * h * w, h * w, w, 1, r1, r2, pad, k, md, s1, s2, corr_type_multiply, stream);
}

int ref_correlation_backward(const float *in1, const float *in2, const float *grad_out, float *grad_in1, float *grad_in2, float *scratch,
                             int b, int c, int h, int w, int pad, int k, int md, int s1, int s2, int corr_type_multiply, cudaStream_t stream) {
    int oc, oh, ow;
    ref_correlation_out_shape(h, w, pad, k, md, s1, s2, &oc, &oh, &ow);
    const long long rn = (long long)b * (h + 2 * pad) * (w + 2 * pad) * c;
    float *r1 = scratch, *r2 = scratch + rn;
    RCK(cudaMemsetAsync(r1, 0, 2 * rn * sizeof(float), stream));
    RCK(cudaMemsetAsync(grad_in1, 0, (size_t)b * c * h * w * sizeof(float), stream));
    RCK(cudaMemsetAsync(grad_in2, 0, (size_t)b * c * h * w * sizeof(float), stream));
    return Correlation_backward_cuda_kernel((float *)grad_out, b, oc, oh, ow, oc * oh * ow, oh * ow, ow, 1, (float *)in1, c, h, w, c * h * w,
                                            h * w, w, 1, (float *)in2, c * h * w, h * w, w, 1, grad_in1, c * h * w, h * w, w, 1, grad_in2, c,
                                            c * h * w, h * w, w, 1, r1, r2, pad, k, md, s1, s2, corr_type_multiply, stream);
}

// img [b_img, c, ih, iw], flow [b, 2, h, w] -> out [b, c, h, w]      (functions/resample2d.py:15-19: output zero-filled first)
int ref_resample2d_forward(const float *img, const float *flow, float *out, int b, int c, int ih, int iw, int h, int w, int kernel_size,
                           cudaStream_t stream) {
    THCState st = {stream, 0};
    THCudaTensor t1 = tensor4(img, b, c, ih, iw), t2 = tensor4(flow, b, 2, h, w), to = tensor4(out, b, c, h, w);
    RCK(cudaMemsetAsync(out, 0, (size_t)b * c * h * w * sizeof(float), stream));
    Resample2d_kernel_forward(&st, &t1, &t2, &to, kernel_size);
    return st.last_error == 0;
}

int ref_resample2d_backward(const float *img, const float *flow, const float *grad_out, float *grad_img, float *grad_flow, int b, int c,
                            int ih, int iw, int h, int w, int kernel_size, cudaStream_t stream) {
    THCState st = {stream, 0};
    THCudaTensor t1 = tensor4(img, b, c, ih, iw), t2 = tensor4(flow, b, 2, h, w), tg = tensor4(grad_out, b, c, h, w);
    THCudaTensor g1 = tensor4(grad_img, b, c, ih, iw), g2 = tensor4(grad_flow, b, 2, h, w);
    RCK(cudaMemsetAsync(grad_img, 0, (size_t)b * c * ih * iw * sizeof(float), stream));
    RCK(cudaMemsetAsync(grad_flow, 0, (size_t)b * 2 * h * w * sizeof(float), stream));
    Resample2d_kernel_backward(&st, &t1, &t2, &tg, &g1, &g2, kernel_size);
    return st.last_error == 0;
}

int ref_channelnorm_forward(const float *x, float *out, int b, int c, int h, int w, int norm_deg, cudaStream_t stream) {
    THCState st = {stream, 0};
    THCudaTensor t1 = tensor4(x, b, c, h, w), to = tensor4(out, b, 1, h, w);
    RCK(cudaMemsetAsync(out, 0, (size_t)b * h * w * sizeof(float), stream));
    ChannelNorm_kernel_forward(&st, &t1, &to, norm_deg);
    return st.last_error == 0;
}

int ref_channelnorm_backward(const float *x, const float *out, const float *grad_out, float *grad_x, int b, int c, int h, int w,
                             int norm_deg, cudaStream_t stream) {
    THCState st = {stream, 0};
    THCudaTensor t1 = tensor4(x, b, c, h, w), to = tensor4(out, b, 1, h, w), tg = tensor4(grad_out, b, 1, h, w), g1 = tensor4(grad_x, b, c, h, w);
    RCK(cudaMemsetAsync(grad_x, 0, (size_t)b * c * h * w * sizeof(float), stream));
    ChannelNorm_kernel_backward(&st, &t1, &to, &tg, &g1, norm_deg);
    return st.last_error == 0;
}

}  // extern "C"
