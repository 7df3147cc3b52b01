// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 on one B200 (registers only, 16 independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, float a, float b, int iters) {
    float2 c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 bb = make_float2(b, b * 1.0001f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) { c[i].x = fmaf(a, bb.x, c[i].x); c[i].y = fmaf(a, bb.y, c[i].y); }
            else c[i] = __ffma2_rn(make_float2(a, a), bb, c[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i].x + c[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *o; cudaMalloc(&o, 148 * 4 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 512>>>(o, 1.0001f, 0.9999f, iters); else k<1><<<148 * 2, 512>>>(o, 1.0001f, 0.9999f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fl = 2.0 * 32 * iters * 148.0 * 2 * 512;
            printf("%s  %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, fl / ms / 1e9);
        }
    return 0;
}
