// Per-kernel-class device timing with CUDA events on the launch stream (bench.py's roofline leg).
// Off by default: when off, VvProfScope costs one branch.  Never enabled inside a timed region that reports `value`.
#include <vector>

#include "common.h"

namespace {
struct Rec {
    cudaEvent_t a, b;
    int cls;
};
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
double g_flops[VV_PROF_CLASSES];
long long g_launches[VV_PROF_CLASSES];

cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

bool vv_prof_on() { return g_on; }

VvProfScope::VvProfScope(int cls, double flops, cudaStream_t st) : idx(-1), st(st) {
    if (!g_on) return;
    Rec r;
    r.a = get_event(); r.b = get_event(); r.cls = cls;
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    idx = (int)g_recs.size() - 1;
    g_flops[cls] += flops;
    g_launches[cls]++;
}
VvProfScope::~VvProfScope() {
    if (idx >= 0) cudaEventRecord(g_recs[idx].b, st);
}

extern "C" int vecvad_profile_begin(void) {
    for (auto &r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
    g_recs.clear();
    for (int i = 0; i < VV_PROF_CLASSES; i++) { g_flops[i] = 0; g_launches[i] = 0; }
    g_on = true;
    return 0;
}

extern "C" int vecvad_profile_end(double *ms, double *flops, int64_t *launches, int n_classes) {
    g_on = false;
    VV_REQUIRE(ms && flops && launches && n_classes >= VV_PROF_CLASSES, "profile_end: need room for %d classes", VV_PROF_CLASSES);
    VV_CK(cudaDeviceSynchronize());
    for (int i = 0; i < VV_PROF_CLASSES; i++) { ms[i] = 0; flops[i] = g_flops[i]; launches[i] = g_launches[i]; }
    for (auto &r : g_recs) {
        float t = 0.f;
        VV_CK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cls] += t;
    }
    return 0;
}
