// tcgen05 (kind::tf32) implicit-GEMM tiles -- placeholder until the TMA/TMEM kernels land.
#include "common.h"
bool vv_igemm_tc_supported(const VvIGemm &) { return false; }
bool vv_wgrad_tc_supported(const VvWGrad &) { return false; }
int vv_launch_igemm_tc(const VvIGemm &, cudaStream_t) { return vv_set_err(-3, "tcgen05 igemm not built"); }
int vv_launch_wgrad_tc(const VvWGrad &, cudaStream_t) { return vv_set_err(-3, "tcgen05 wgrad not built"); }
