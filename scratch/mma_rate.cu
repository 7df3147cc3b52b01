// microbench: cycles per tcgen05.mma for kind::tf32 / kind::f16, K-major SW128 operands in smem, various N, single-lane vs warp-uniform issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vec_vad_b200/csrc/tc_common.cuh"

__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int MODE>   // 0: tf32 single lane, 1: tf32 whole warp + elect, 2: bf16 single lane, 3: tf32 warp+elect walking distinct tiles
__global__ void __launch_bounds__(128) k(int N, int M, int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (MODE == 3 ? 200 * 1024 : (128 + 256) * 128) / 4; i += 128) ((float *)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem = slot;
    uint32_t idesc;
    if (MODE == 2) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    else idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint64_t da = smem_desc_k_sw128(smem_u32(smem)), db = smem_desc_k_sw128(smem_u32(smem + 128 * 128));
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        if (MODE == 3) {
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
                const uint32_t off = (uint32_t)((i % 8) * 20 * 1024);          // 8 distinct 20 KB regions: A at +0, B at +16 KB
                const uint64_t da2 = smem_desc_k_sw128(smem_u32(smem) + off), db2 = smem_desc_k_sw128(smem_u32(smem) + off + 16384);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (elect_one()) tc_mma_tf32(tmem, da2 + 2 * k, db2 + 2 * k, idesc, 1u);
            }
            if (elect_one()) tc_commit(&bar);
            __syncwarp();
        } else if (MODE == 1) {
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (elect_one()) tc_mma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
            }
            if (elect_one()) tc_commit(&bar);
            __syncwarp();
        } else if (threadIdx.x == 0) {
            t0 = clock64();
            for (int i = 0; i < iters; i++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (MODE == 2) mma_f16(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                    else tc_mma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                }
            }
            tc_commit(&bar);
        }
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    long long *d; cudaMalloc(&d, 148 * 8);
    long long h[148];
    const int iters = 2000;
    int smem = (128 + 256) * 128 + 2048;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int M : {128}) for (int N : {32, 64, 128, 256}) {
        for (int mode = 0; mode < 4; mode++) {
            if (mode == 3 && N > 32) continue;
            for (int grid : {1, 148}) {
                if (mode == 0) k<0><<<grid, 128, smem>>>(N, M, iters, d);
                if (mode == 1) k<1><<<grid, 128, smem>>>(N, M, iters, d);
                if (mode == 2) k<2><<<grid, 128, smem>>>(N, M, iters, d);
                if (mode == 3) k<3><<<grid, 128, 210 * 1024>>>(N, M, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                double cyc = (double)h[0] / (iters * 4);
                int K = mode == 2 ? 16 : 8;
                printf("M=%d N=%3d mode=%d(%s) grid=%3d: %.1f cycles/MMA  -> %.0f MAC/cycle/SM  (%s)\n", M, N, mode,
                       mode == 0 ? "tf32 lane0" : mode == 1 ? "tf32 warp+elect" : mode == 2 ? "bf16 lane0" : "tf32 warp+elect distinct tiles", grid, cyc, (double)M * N * K / cyc, cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
