// Which tiled-TMA configurations are legal on this part?  usage: tma_probe <variant>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../vec_vad_b200/csrc/tc_common.cuh"

__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, int c0, int c1, int c2, int c3, int nfloats, int dst_off) {
    extern __shared__ uint8_t raw[];
    float *sm = (float *)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = -7.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, nfloats * 4);
        tma_load_4d(sm + dst_off, &tm, &bar, c0, c1, c2, c3);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = sm[dst_off + i];
}

int main(int argc, char **argv) {
    int v = argc > 1 ? atoi(argv[1]) : 1;
    const int W = 64, H = 12, C = 32, N = 2;
    std::vector<float> h((size_t)W * H * C * N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 1000);
    float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1 << 20);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cuuint64_t dims[4] = {W, H, C, N}; cuuint64_t strides[3] = {W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {32, 1, 8, 1}, estr[4] = {1, 1, 1, 1};
    int c0 = 0, c1 = 2, c2 = 8, c3 = 1, dst_off = 0;
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    const char *name = "";
    switch (v) {
        case 1: name = "none 32x1x8 in-bounds"; break;
        case 2: name = "none 32x1x8 x=-10"; c0 = -10; break;
        case 3: name = "none 32x1x8 row=-5 (fully OOB)"; c1 = -5; break;
        case 4: name = "none 52x1x8 in-bounds"; box[0] = 52; break;
        case 5: name = "none 52x21x8 estr1 y=-4"; box[0] = 52; box[1] = 21; c1 = -4; break;
        case 6: name = "none 52x42x8 estr2 y=-4"; box[0] = 52; box[1] = 42; estr[1] = 2; c1 = -4; break;
        case 7: name = "none 64x21x8"; box[0] = 64; box[1] = 21; c1 = -4; break;
        case 8: name = "none 32x1x8 dst+2048B"; dst_off = 512; break;
        case 9: name = "none 52x21x8 dst+2048B x=-10"; box[0] = 52; box[1] = 21; c0 = -10; c1 = -4; dst_off = 512; break;
        case 10: name = "none 32x21x8 estr2"; box[1] = 42; estr[1] = 2; c1 = -4; break;
        case 11: name = "none 48x21x8"; box[0] = 48; box[1] = 21; break;
        case 12: name = "none 56x21x8"; box[0] = 56; box[1] = 21; break;
    }
    int nfl = box[0] * ((box[1] + estr[1] - 1) / estr[1]) * box[2] * box[3];
    CUtensorMap tm;
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("V%d %-36s encode failed %d\n", v, name, (int)r); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<1, 128, 100 * 1024>>>(tm, o, c0, c1, c2, c3, nfl, dst_off);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> res(nfl);
    if (e == cudaSuccess) cudaMemcpy(res.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
    // expected first in-bounds element
    double sum = 0; for (float x : res) sum += x;
    printf("V%d %-36s -> %s  nfloats %d  sum %.0f  first %.0f\n", v, name, cudaGetErrorString(e), nfl, sum, nfl ? res[0] : -1.f);
    return 0;
}
