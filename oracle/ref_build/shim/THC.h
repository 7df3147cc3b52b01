/* TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the legacy THC headers, just enough for the reference's
 * Resample2d_kernel.cu / ChannelNorm_kernel.cu (FlowNet2_src/models/components/ops/{resample2d,channelnorm}/src) to compile
 * UNMODIFIED, from where they lie under /root/reference, with nvcc 12.9 (torch 2.x no longer ships THC).  Only the members
 * those two files touch exist: ->size[i], ->stride[i], THCudaTensor_nElement, THCudaTensor_data, THCState_getCurrentStream,
 * THCudaCheck.  Nothing in vec_vad_b200/ includes or links this. */
#ifndef VECVAD_ORACLE_THC_SHIM_H
#define VECVAD_ORACLE_THC_SHIM_H
#include <cuda_runtime.h>
#include <stdio.h>

typedef struct THCState {
    cudaStream_t stream;
    int last_error;            /* set by THCudaCheck: the harness reads it instead of aborting */
} THCState;

typedef struct THCudaTensor {
    float *data;
    long size[4];
    long stride[4];
} THCudaTensor;

static inline long THCudaTensor_nElement(THCState *s, THCudaTensor *t) {
    (void)s;
    return t->size[0] * t->size[1] * t->size[2] * t->size[3];
}
static inline float *THCudaTensor_data(THCState *s, THCudaTensor *t) {
    (void)s;
    return t->data;
}
static inline cudaStream_t THCState_getCurrentStream(THCState *s) { return s->stream; }

#define THCudaCheck(expr)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            printf("THCudaCheck (shim): %s\n", cudaGetErrorString(e__));                    \
            state->last_error = (int)e__;                                                   \
        }                                                                                   \
    } while (0)
#endif
