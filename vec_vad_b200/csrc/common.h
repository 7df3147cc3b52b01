// Shared declarations for libvecvad.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "vecvad.h"

int vv_set_err(int code, const char *fmt, ...);

#define VV_CK(call)                                                                                        \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return vv_set_err(-2, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
    } while (0)
extern unsigned long long g_vv_launches;   // kernels launched by this library (bench.py reports it as gpu_launches)
#define VV_CKL()                      \
    do {                              \
        g_vv_launches++;              \
        VV_CK(cudaGetLastError());    \
    } while (0)
#define VV_REQUIRE(cond, ...)                          \
    do {                                               \
        if (!(cond)) return vv_set_err(-1, __VA_ARGS__); \
    } while (0)

static inline int vv_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the train step is launched through vv_launch(): with the programmatic stream
// serialization attribute the NEXT kernel's CTAs are scheduled as soon as all CTAs of this one have passed vv_pdl_wait() (or
// exited) and free SM slots exist, so launch latency, barrier / TMEM set-up and tensor-map fetches of kernel i+1 overlap the
// tail of kernel i.  Contract: a kernel touches NO global memory another kernel writes (reads or writes) before vv_pdl_wait(),
// which blocks until every preceding kernel in the stream has completed and its writes are visible.  Without the attribute
// (VECVAD_PDL=0) both instructions are no-ops.
// ---------------------------------------------------------------------------------------------
bool vv_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void vv_pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t vv_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (vv_pdl_enabled()) { cfg.attrs = at; cfg.numAttrs = 1; }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM problem descriptions shared by the SIMT (fp32) and tcgen05 (tf32) tile kernels.
// Activations are NHWC fp32, "grouped": [G][B*H*W][ld] with a uniform group stride, so the G
// independent UNets of a net run in one launch (blockIdx.z = g).
// ---------------------------------------------------------------------------------------------
struct VvTaps {
    int n;
    int dy[9];
    int dx[9];
};

// out[m, n] = bias[n] + sum_t sum_k A[shift(m, t), k] * Wt[t][n][k]
struct VvIGemm {
    const float *A;     long long a_gs;  int lda, a_coff, a_s2d;   // a_s2d: A is the space-to-depth view of a [B,2H,2W,Kt/4] tensor
    int Kt;                                                        // K per tap (multiple of 16)
    int B, H, W;                                                   // logical pixel grid
    const float *Wt;    long long w_gs;                            // [ntaps][N][Kt]
    VvTaps taps;
    int N;                                                         // multiple of 16
    float *O;           long long o_gs;  int ldo, o_coff, o_d2s;   // o_d2s: N = 4*Co, pixel-shuffled into a [B,2H,2W,*] tensor
    const float *bias;  long long bias_gs;                         // nullable; indexed by n (or co when o_d2s)
    double *stats;      long long stats_gs;                        // nullable; [2][N] column sum / sum of squares
    int G;
    int ab_f16;                                                    // A and Wt hold fp16 (lda / a_coff / a_gs / w_gs / Kt in elements); tcgen05 tiles only
    int o_f16;                                                     // O holds fp16 (ldo / o_coff / o_gs in elements): the transposed conv writing the concat half
    // split output (input gradient of the first conv of an up-block): columns >= o_split go, as fp16, to O2 (the transposed
    // conv's output gradient) instead of O; 0 = no split
    int o_split;
    void *O2;           long long o2_gs; int ldo2;
    int rev;                                                       // tcgen05 tiles: walk the pixel tiles from the last to the first (serpentine order for L2 reuse)
};

// dW[t][n][k] += sum_m Gd[m, n] * A[shift(m, t), k]
struct VvWGrad {
    const float *A;     long long a_gs;  int lda, a_coff;
    int Kt;
    int B, H, W;
    const float *Gd;    long long g_gs;  int ldg, g_coff, g_s2d;   // g_s2d: Gd is the space-to-depth view of a [B,2H,2W,N/4] tensor
    int N;
    VvTaps taps;
    float *dW;          long long dw_gs;                           // [ntaps][N][Kt], pre-zeroed, atomically accumulated
    int G;
    int ab_f16;                                                    // A and Gd hold fp16 (strides / offsets in elements); tcgen05 tiles only
};

// ---- per-kernel-class event timing (prof.cu)
enum { VV_PROF_IGEMM_TC = 0, VV_PROF_WGRAD_TC, VV_PROF_IGEMM_SIMT, VV_PROF_WGRAD_SIMT, VV_PROF_BN, VV_PROF_OTHER, VV_PROF_CLASSES };
bool vv_prof_on();
struct VvProfScope {
    int idx;
    cudaStream_t st;
    VvProfScope(int cls, double flops, cudaStream_t st);
    ~VvProfScope();
};

int vv_launch_igemm_simt(const VvIGemm &p, cudaStream_t st);
int vv_launch_wgrad_simt(const VvWGrad &p, cudaStream_t st);
// tcgen05 tiles (kind::tf32, or kind::f16 when ab_f16): which problems they accept at all (tc_support.cu) ...
bool vv_igemm_tc_supported(const VvIGemm &p);
bool vv_wgrad_tc_supported(const VvWGrad &p);
// ... flattened-sequence 3x3 conv tiles (igemm_flat.cu): one activation box serves all nine taps.  _shape_ok: what the kernel can
// run; _supported: where the net engine prefers it (full-width rows, VECVAD_FLAT != 0)
bool vv_igemm_flat_shape_ok(const VvIGemm &p);
bool vv_igemm_flat_supported(const VvIGemm &p);
int vv_launch_igemm_flat(const VvIGemm &p, cudaStream_t st);
// ... pair tiles (igemm_tc3.cu): two pixel tiles share every weight tile, two MMA warps, two epilogue sets
bool vv_igemm_tc3_supported(const VvIGemm &p);
int vv_launch_igemm_tc3(const VvIGemm &p, cudaStream_t st);
// ... flattened-sequence weight-gradient tiles (wgrad_flat.cu): one MMA per K-step covers all nine taps
bool vv_wgrad_flat_shape_ok(const VvWGrad &p);
bool vv_wgrad_flat_supported(const VvWGrad &p);
int vv_launch_wgrad_flat(const VvWGrad &p, cudaStream_t st);
// ... tap-reuse weight-gradient tiles (wgrad_tc2.cu)
bool vv_wgrad_tc2_supported(const VvWGrad &p);
int vv_launch_wgrad_tc2(const VvWGrad &p, cudaStream_t st);
