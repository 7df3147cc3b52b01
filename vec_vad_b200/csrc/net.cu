// The completion-UNet set engine: plans the workspace, sequences the kernels of one forward /
// backward over G independent UNets (grouped launches), behind the C ABI of include/vecvad.h.
//
// Reference data flow restated: model/unet.py:172-267 (SelfCompleteNet4.forward), :410-556
// (SelfCompleteNetFull.forward), :619-652 (SelfCompleteNet1raw1of.forward); one UNet = inc -> down x3
// -> up x3 -> outc (model/unet.py:187-196); loss + backward = train.py:385-402.
#include <new>
#include <stdlib.h>

#include "unet_kernels.h"

static thread_local char g_err[768] = "";

int vv_set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *vecvad_last_error(void) { return g_err; }
extern "C" int vecvad_abi_version(void) { return VECVAD_ABI_VERSION; }
unsigned long long g_vv_launches = 0;
bool vv_pdl_enabled() {                      // VECVAD_PDL=0: plain stream-ordered launches (A/B baseline)
    static int on = -1;
    if (on < 0) { const char *e = getenv("VECVAD_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}
extern "C" uint64_t vecvad_launch_count(void) { return g_vv_launches; }

namespace {

struct View {           // grouped NHWC view: element (g, m, c) at p + g*gs + m*ld + coff + c   (element units: fp32, or fp16 for the
    void *p;            // activation buffers of the fp16-operand mode)
    long long gs;
    int ld, coff, C, H;
};

constexpr int NU = VECVAD_N_UNITS;
constexpr int NT = VECVAD_N_UPS;

}  // namespace

#define VV_NEV 48
struct vecvad_net {
    vecvad_net_config cfg;
    int G, F, S, T, cin_real, cinp;
    int serp;                                // serpentine row / tile order through consecutive kernels (VECVAD_SERPENTINE, default on)
    int f16;                                 // use_tensor_cores == 2: activations, dZ and re-laid-out weights are fp16 (operands of kind::f16 tiles)
    float lscale, lscale_user;               // power-of-two loss scale of the fp16 gradient operands (1 in the other modes); user override (<= 0: automatic)
    VvIntG slot, erase, outc, isflow, tidx, oslot;
    int n_raw_out, n_of_out, n_raw_tot, n_of_tot;
    // unit geometry: conv unit u maps C -> N at resolution H
    int uC[NU], uCp[NU], uN[NU], uH[NU];
    int tCi[NT], tCo[NT], tH[NT];          // transposed convs: Ci -> Co, input resolution tH
    // bound buffers
    float *params, *grads, *running;
    char *ws;
    int64_t ws_bytes;
    int maxB;
    // workspace sub-buffers (float offsets resolved to pointers at bind time)
    void *X0, *A[NU], *CAT[3], *PL[3], *X4, *UU[3], *Z[NU];   // activations and raw conv outputs: fp32, or fp16 in the fp16-operand mode
    void *dCAT[3], *GA, *GB;                                  // output gradients (dY): fp32, or (loss-scaled) fp16 in the fp16-operand mode
    float *DOUT;                                              // staged d loss / d out [m][4]: fp32
    void *dUP[3], *HZ[4];                                     // fp16 mode only: transposed-conv output gradients, ring of dZ operand buffers
    void *Wf[NU], *Wd[NU], *tWf[NT], *tWd[NT];                // re-laid-out weights: fp32 or fp16
    float *vec[NU], *save[NU], *tvec[NT];
    float *dWf[NU], *tdW[NT];
    double *stats[NU], *bsums[NU];
    char *zero_fwd;  size_t zero_fwd_bytes;   // BN statistics accumulators
    char *zero_bwd;  size_t zero_bwd_bytes;   // weight-gradient accumulators + BN backward sums
    // state of the last forward
    int lastB, last_training, have_dout;
    // weight-gradient tiles run on a side stream (they are off the backward critical path: nothing downstream reads dW)
    cudaStream_t wg_stream;
    cudaEvent_t ev[VV_NEV];
    int ev_next, use_side;
    // gradient phases (vecvad_net_grad_phase_*): the parameter gradients of a slot become final in three contiguous ranges, in
    // this order: [unit 8, end) after the decoder, [unit 6, unit 8) after the deepest encoder block, [0, unit 6) at the end
    cudaEvent_t ev_phase[3];
    int have_phase_ev, phases_recorded;
    // the backward's zero-fills (weight-gradient accumulators, BN backward sums, the gradient slots) issued on the side stream by the
    // training forward, off the critical path; ev_zero: recorded behind them
    cudaEvent_t ev_zero;
    int have_zero_ev, bwd_zeroed;
    int defer_join;                          // vecvad_net_defer_join: the backward leaves the side stream unjoined (caller waits for phase 2)
};

namespace {

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// Walks the workspace layout; with base == nullptr it only measures.
long long layout(vecvad_net *n, int B, char *base) {
    long long off = 0;
    auto take = [&](long long bytes) -> char * {
        char *p = base ? base + off : nullptr;
        off = align_up(off + bytes, 256);
        return p;
    };
    const int G = n->G, F = n->F, S = n->S;
    const long long es = n->f16 ? 2 : 4;                       // bytes per activation / operand element
    auto M = [&](int H) { return (long long)B * H * H; };
    auto fl = [&](long long count) { return (float *)take(count * (long long)sizeof(float)); };
    auto act = [&](long long count) { return (void *)take(count * es); };
    n->X0 = act(G * M(S) * n->cinp);
    for (int u = 0; u < NU; u++) n->Z[u] = act(G * M(n->uH[u]) * n->uN[u]);
    for (int u = 0; u < NU; u += 2) n->A[u] = act(G * M(n->uH[u]) * n->uN[u]);
    // CAT[k]: concat input of up-block k+1:  k=0 @S/4 (8F), k=1 @S/2 (4F), k=2 @S (2F)
    for (int k = 0; k < 3; k++) {
        int H = S >> (2 - k), C = F << (3 - k);
        n->CAT[k] = act(G * M(H) * C);
        n->dCAT[k] = act(G * M(H) * C);
        n->dUP[k] = n->f16 ? take(G * M(H) * (C / 2) * 2) : nullptr;
    }
    // PL[k]: pooled input of down-block k+1: k=0 @S/2 (F), k=1 @S/4 (2F), k=2 @S/8 (4F)
    for (int k = 0; k < 3; k++) n->PL[k] = act(G * M(S >> (k + 1)) * (F << k));
    n->X4 = act(G * M(S >> 3) * 8 * F);
    for (int k = 0; k < 3; k++) n->UU[k] = act(G * M(S >> (2 - k)) * (F << (2 - k)));   // U1 @S/4 4F, U2 @S/2 2F, U3 @S F
    n->GA = act(G * M(S) * F);
    n->GB = act(G * M(S) * F);
    for (int i = 0; i < 4; i++) n->HZ[i] = n->f16 ? take(G * M(S) * F * 2) : nullptr;
    n->DOUT = fl(G * M(S) * 4);
    for (int u = 0; u < NU; u++) {
        n->Wf[u] = act((long long)G * 9 * n->uN[u] * n->uCp[u]);
        n->Wd[u] = act((long long)G * 9 * n->uN[u] * n->uCp[u]);
        n->vec[u] = fl((long long)G * 3 * n->uN[u]);
        n->save[u] = fl((long long)G * 4 * n->uN[u]);
    }
    for (int k = 0; k < NT; k++) {
        n->tWf[k] = act((long long)G * 16 * n->tCo[k] * n->tCi[k]);
        n->tWd[k] = act((long long)G * 16 * n->tCo[k] * n->tCi[k]);
        n->tvec[k] = fl((long long)G * n->tCo[k]);
    }
    long long z0 = off;
    for (int u = 0; u < NU; u++) n->stats[u] = (double *)take((long long)G * 2 * n->uN[u] * sizeof(double));
    n->zero_fwd = base ? base + z0 : nullptr;
    n->zero_fwd_bytes = (size_t)(off - z0);
    long long z1 = off;
    for (int u = 0; u < NU; u++) {
        n->dWf[u] = fl((long long)G * 9 * n->uN[u] * n->uCp[u]);
        n->bsums[u] = (double *)take((long long)G * 2 * n->uN[u] * sizeof(double));
    }
    for (int k = 0; k < NT; k++) n->tdW[k] = fl((long long)G * 16 * n->tCo[k] * n->tCi[k]);
    n->zero_bwd = base ? base + z1 : nullptr;
    n->zero_bwd_bytes = (size_t)(off - z1);
    return off;
}

VvTaps taps3x3(int sign) {
    VvTaps t;
    t.n = 9;
    for (int k = 0; k < 9; k++) { t.dy[k] = sign * (k / 3 - 1); t.dx[k] = sign * (k % 3 - 1); }
    return t;
}
VvTaps taps2x2(int sign) {
    VvTaps t;
    t.n = 4;
    for (int k = 0; k < 4; k++) { t.dy[k] = sign * (k >> 1); t.dx[k] = sign * (k & 1); }
    for (int k = 4; k < 9; k++) { t.dy[k] = 0; t.dx[k] = 0; }
    return t;
}

// algorithmic FLOPs of one contraction launch: 2 * pixels * N * K * taps over all groups
double igemm_flops(int B, int H, int W, int N, int K, int taps, int G) { return 2.0 * B * H * W * (double)N * K * taps * G; }

// k_real: real K per tap where the operand is channel-padded (first conv); flops > 0: the algorithmic count where the launch computes
// structural zeros (the stride-2 transposed conv runs 16 tap-blocks for the 9 taps of its 3x3 kernel)
int run_igemm(bool want_tc, const VvIGemm &p, cudaStream_t st, int k_real = 0, double flops = 0.0) {
    const bool tc = want_tc && vv_igemm_tc_supported(p);
    if (p.ab_f16 && !(tc && (vv_igemm_flat_supported(p) || vv_igemm_tc3_supported(p))))
        return vv_set_err(-3, "fp16-operand contraction %dx%d Kt=%d N=%d is not covered by the tcgen05 tiles (no fp32 fallback reads fp16)", p.H, p.W, p.Kt, p.N);
    VvProfScope ps(tc ? VV_PROF_IGEMM_TC : VV_PROF_IGEMM_SIMT, flops > 0.0 ? flops : igemm_flops(p.B, p.H, p.W, p.N, k_real ? k_real : p.Kt, p.taps.n, p.G), st);
    if (tc && vv_igemm_flat_supported(p)) return vv_launch_igemm_flat(p, st);
    if (tc && vv_igemm_tc3_supported(p)) return vv_launch_igemm_tc3(p, st);
    return vv_launch_igemm_simt(p, st);
}
int run_wgrad(bool want_tc, const VvWGrad &p, cudaStream_t st, int k_real = 0, double flops = 0.0) {
    const bool tc = want_tc && vv_wgrad_tc_supported(p);
    if (p.ab_f16 && !(tc && (vv_wgrad_flat_supported(p) || vv_wgrad_tc2_supported(p))))
        return vv_set_err(-3, "fp16-operand weight gradient %dx%d Kt=%d N=%d is not covered by the tcgen05 tiles", p.H, p.W, p.Kt, p.N);
    VvProfScope ps(tc ? VV_PROF_WGRAD_TC : VV_PROF_WGRAD_SIMT, flops > 0.0 ? flops : igemm_flops(p.B, p.H, p.W, p.N, k_real ? k_real : p.Kt, p.taps.n, p.G), st);
    if (tc && vv_wgrad_flat_supported(p)) return vv_launch_wgrad_flat(p, st);
    if (tc && vv_wgrad_tc2_supported(p)) return vv_launch_wgrad_tc2(p, st);
    return vv_launch_wgrad_simt(p, st);
}

struct Flow {   // buffer wiring of one forward for batch B
    View in[NU], y[NU];
    int pool[NU];      // index into PL or -1
    View tin[NT], tout[NT];
};

View mk(void *p, int B, int H, int ld, int coff, int C) {
    View v;
    v.p = p; v.gs = (long long)B * H * H * ld; v.ld = ld; v.coff = coff; v.C = C; v.H = H;
    return v;
}

void wire(const vecvad_net *n, int B, Flow &f) {
    const int F = n->F, S = n->S;
    for (int u = 0; u < NU; u++) f.pool[u] = -1;
    // encoder
    f.in[0] = mk(n->X0, B, S, n->cinp, 0, n->cinp);
    f.y[0] = mk(n->A[0], B, S, F, 0, F);
    f.in[1] = f.y[0];
    f.y[1] = mk(n->CAT[2], B, S, 2 * F, 0, F);            f.pool[1] = 0;
    f.in[2] = mk(n->PL[0], B, S / 2, F, 0, F);
    f.y[2] = mk(n->A[2], B, S / 2, 2 * F, 0, 2 * F);
    f.in[3] = f.y[2];
    f.y[3] = mk(n->CAT[1], B, S / 2, 4 * F, 0, 2 * F);    f.pool[3] = 1;
    f.in[4] = mk(n->PL[1], B, S / 4, 2 * F, 0, 2 * F);
    f.y[4] = mk(n->A[4], B, S / 4, 4 * F, 0, 4 * F);
    f.in[5] = f.y[4];
    f.y[5] = mk(n->CAT[0], B, S / 4, 8 * F, 0, 4 * F);    f.pool[5] = 2;
    f.in[6] = mk(n->PL[2], B, S / 8, 4 * F, 0, 4 * F);
    f.y[6] = mk(n->A[6], B, S / 8, 8 * F, 0, 8 * F);
    f.in[7] = f.y[6];
    f.y[7] = mk(n->X4, B, S / 8, 8 * F, 0, 8 * F);
    // decoder: up-block k (0..2): convT(tin -> second half of CAT[k]) ; conv units 8+2k, 9+2k
    for (int k = 0; k < 3; k++) {
        int H = S >> (2 - k), C = F << (3 - k);           // concat resolution / channels
        f.tin[k] = (k == 0) ? f.y[7] : f.y[7 + 2 * k];    // X4, U1, U2
        f.tout[k] = mk(n->CAT[k], B, H, C, C / 2, C / 2);
        int u = 8 + 2 * k;
        f.in[u] = mk(n->CAT[k], B, H, C, 0, C);
        f.y[u] = mk(n->A[u], B, H, C / 2, 0, C / 2);
        f.in[u + 1] = f.y[u];
        f.y[u + 1] = mk(n->UU[k], B, H, C / 2, 0, C / 2);
    }
}

}  // namespace

// dispatchers shared with single_ops.cu
int vv_run_igemm(bool want_tc, const VvIGemm &p, cudaStream_t st) { return run_igemm(want_tc, p, st); }
int vv_run_wgrad(bool want_tc, const VvWGrad &p, cudaStream_t st) { return run_wgrad(want_tc, p, st); }
VvTaps vv_taps3x3(int sign) { return taps3x3(sign); }
VvTaps vv_taps2x2(int sign) { return taps2x2(sign); }

extern "C" int vecvad_net_create(const vecvad_net_config *cfg, vecvad_net **out) {
    VV_REQUIRE(cfg && out, "net_create: null argument");
    VV_REQUIRE(cfg->n_unets >= 1 && cfg->n_unets <= VECVAD_MAX_UNETS, "net_create: n_unets=%d out of range", cfg->n_unets);
    VV_REQUIRE(cfg->features_root >= 16 && cfg->features_root % 16 == 0, "net_create: features_root=%d must be a multiple of 16",
               cfg->features_root);
    VV_REQUIRE(cfg->patch >= 16 && cfg->patch % 16 == 0, "net_create: patch=%d must be a multiple of 16", cfg->patch);
    VV_REQUIRE(cfg->tot_raw_num >= 2 && cfg->tot_raw_num <= 10, "net_create: tot_raw_num=%d unsupported", cfg->tot_raw_num);
    vecvad_net *n = new (std::nothrow) vecvad_net();
    VV_REQUIRE(n, "net_create: out of host memory");
    memset(n, 0, sizeof(*n));
    n->cfg = *cfg;
    n->G = cfg->n_unets; n->F = cfg->features_root; n->S = cfg->patch; n->T = cfg->tot_raw_num;
    n->cin_real = 3 * (cfg->padding ? n->T : n->T - 1);
    n->cinp = (n->cin_real + 15) / 16 * 16;
    if (cfg->use_tensor_cores) n->cinp = (n->cin_real + 31) / 32 * 32;   // tcgen05 tiles use 32-channel K slabs (128 bytes of tf32, 64 of fp16)
    n->f16 = cfg->use_tensor_cores == 2;
    {
        const char *e = getenv("VECVAD_SERPENTINE");
        n->serp = !(e && e[0] == '0');
    }
    n->lscale = 1.f; n->lscale_user = 0.f;
    int max_raw = -1, max_of = -1;
    for (int g = 0; g < n->G; g++) {
        n->slot.v[g] = cfg->param_slot[g]; n->erase.v[g] = cfg->erase_frame[g]; n->outc.v[g] = cfg->out_channels[g];
        n->isflow.v[g] = cfg->target_is_flow[g]; n->tidx.v[g] = cfg->target_index[g]; n->oslot.v[g] = cfg->out_slot[g];
        if (!(cfg->erase_frame[g] >= 0 && cfg->erase_frame[g] < n->T) || !(cfg->out_channels[g] == (cfg->target_is_flow[g] ? 2 : 3)) ||
            cfg->param_slot[g] < 0 || cfg->out_slot[g] < 0 || cfg->target_index[g] < 0) {
            delete n;
            return vv_set_err(-1, "net_create: bad per-UNet configuration at g=%d", g);
        }
        if (cfg->target_is_flow[g]) max_of = cfg->out_slot[g] > max_of ? cfg->out_slot[g] : max_of;
        else max_raw = cfg->out_slot[g] > max_raw ? cfg->out_slot[g] : max_raw;
    }
    n->n_raw_out = max_raw + 1; n->n_of_out = max_of + 1;
    n->n_raw_tot = cfg->n_raw_total > 0 ? cfg->n_raw_total : n->n_raw_out;
    n->n_of_tot = cfg->n_of_total > 0 ? cfg->n_of_total : n->n_of_out;
    const int F = n->F, S = n->S;
    // conv units: (C -> N @H)
    int C[NU] = {n->cin_real, F, F, 2 * F, 2 * F, 4 * F, 4 * F, 8 * F, 8 * F, 4 * F, 4 * F, 2 * F, 2 * F, F};
    int N[NU] = {F, F, 2 * F, 2 * F, 4 * F, 4 * F, 8 * F, 8 * F, 4 * F, 4 * F, 2 * F, 2 * F, F, F};
    int H[NU] = {S, S, S / 2, S / 2, S / 4, S / 4, S / 8, S / 8, S / 4, S / 4, S / 2, S / 2, S, S};
    for (int u = 0; u < NU; u++) { n->uC[u] = C[u]; n->uCp[u] = (u == 0) ? n->cinp : C[u]; n->uN[u] = N[u]; n->uH[u] = H[u]; }
    for (int k = 0; k < NT; k++) { n->tCi[k] = F << (3 - k); n->tCo[k] = F << (2 - k); n->tH[k] = S >> (3 - k); }
    {
        const char *e = getenv("VECVAD_WGRAD_STREAM");
        n->use_side = !(e && e[0] == '0');
        if (n->use_side) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = least priority: the critical path keeps the SMs
            if (cudaStreamCreateWithPriority(&n->wg_stream, cudaStreamNonBlocking, lo) != cudaSuccess) n->use_side = 0;
            for (int i = 0; n->use_side && i < VV_NEV; i++)
                if (cudaEventCreateWithFlags(&n->ev[i], cudaEventDisableTiming) != cudaSuccess) n->use_side = 0;
        }
    }
    n->have_phase_ev = 1;
    for (int i = 0; i < 3; i++)
        if (cudaEventCreateWithFlags(&n->ev_phase[i], cudaEventDisableTiming) != cudaSuccess) { n->have_phase_ev = 0; break; }
    n->phases_recorded = 0;
    n->have_zero_ev = cudaEventCreateWithFlags(&n->ev_zero, cudaEventDisableTiming) == cudaSuccess;
    n->bwd_zeroed = 0;
    n->defer_join = 0;
    *out = n;
    return 0;
}

extern "C" void vecvad_net_destroy(vecvad_net *net) {
    if (!net) return;
    if (net->have_phase_ev)
        for (int i = 0; i < 3; i++) cudaEventDestroy(net->ev_phase[i]);
    if (net->have_zero_ev) cudaEventDestroy(net->ev_zero);
    if (net->use_side) {
        cudaStreamSynchronize(net->wg_stream);
        for (int i = 0; i < VV_NEV; i++) cudaEventDestroy(net->ev[i]);
        cudaStreamDestroy(net->wg_stream);
    }
    delete net;
}

extern "C" int vecvad_net_workspace_bytes(const vecvad_net *net, int batch, int64_t *bytes) {
    VV_REQUIRE(net && bytes && batch >= 1, "workspace_bytes: bad arguments");
    vecvad_net tmp = *net;
    *bytes = layout(&tmp, batch, nullptr);
    return 0;
}

extern "C" int vecvad_net_bind(vecvad_net *net, float *params, float *grads, float *running_stats, void *workspace,
                               int64_t workspace_bytes, int max_batch) {
    VV_REQUIRE(net && params && workspace && max_batch >= 1, "net_bind: bad arguments");
    VV_REQUIRE(((uintptr_t)workspace) % 256 == 0, "net_bind: workspace must be 256-byte aligned");
    VV_REQUIRE(((uintptr_t)params) % 16 == 0 && ((uintptr_t)grads) % 16 == 0, "net_bind: params/grads must be 16-byte aligned");
    long long need = layout(net, max_batch, (char *)workspace);
    VV_REQUIRE(need <= workspace_bytes, "net_bind: workspace too small (%lld needed, %lld given)", need, (long long)workspace_bytes);
    net->params = params; net->grads = grads; net->running = running_stats;
    net->ws = (char *)workspace; net->ws_bytes = workspace_bytes; net->maxB = max_batch;
    net->lastB = 0; net->last_training = 0; net->have_dout = 0;
    return 0;
}

extern "C" int vecvad_net_forward(vecvad_net *n, const float *x, const float *x_of, int x_of_channels, int batch, int training,
                                  float *raw_out, int raw_out_channels, float *of_out, int of_out_channels, float *sse, float lambda_raw,
                                  float lambda_of, vecvad_stream stream) {
    VV_REQUIRE(n && n->ws, "net_forward: net not bound");
    VV_REQUIRE(x && batch >= 1 && batch <= n->maxB, "net_forward: batch=%d outside 1..%d", batch, n->maxB);
    VV_REQUIRE(n->running, "net_forward: running-statistics buffer not bound");
    VV_REQUIRE(!training || batch * (n->S / 8) * (n->S / 8) > 1, "net_forward: BatchNorm needs more than one value per channel in training");
    if (raw_out) VV_REQUIRE(raw_out_channels >= 3 * n->n_raw_out, "net_forward: raw_out has %d channels, need %d", raw_out_channels, 3 * n->n_raw_out);
    if (n->n_of_out > 0) {
        if (of_out) VV_REQUIRE(of_out_channels >= 2 * n->n_of_out, "net_forward: of_out has %d channels, need %d", of_out_channels, 2 * n->n_of_out);
        if (sse) VV_REQUIRE(x_of != nullptr, "net_forward: x_of required for flow targets");
    }
    for (int g = 0; g < n->G; g++) {
        if (sse && n->isflow.v[g]) VV_REQUIRE(2 * n->tidx.v[g] + 2 <= x_of_channels, "net_forward: flow target %d outside x_of (%d channels)", n->tidx.v[g], x_of_channels);
        if (sse && !n->isflow.v[g]) VV_REQUIRE(n->tidx.v[g] < n->T, "net_forward: raw target %d outside the cube", n->tidx.v[g]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int G = n->G, B = batch, F = n->F, S = n->S;
    const vecvad_net_config &c = n->cfg;
    Flow f;
    wire(n, B, f);
    // 1. weights into GEMM layouts (they change every optimiser step): one launch for all conv units when they tile by 32
    bool batched = true;
    for (int u = 0; u < NU; u++) batched = batched && n->uN[u] % 32 == 0 && n->uCp[u] % 32 == 0;
    cudaStream_t sP = n->use_side ? n->wg_stream : st;
    if (n->use_side) {
        VV_CK(cudaEventRecord(n->ev[VV_NEV - 1], st));
        VV_CK(cudaStreamWaitEvent(sP, n->ev[VV_NEV - 1], 0));
    }
    if (batched) {
        VvPrepAll all;
        memset(&all, 0, sizeof(all));
        all.n = NU;
        for (int u = 0; u < NU; u++) {
            VvPrepUnit &pu = all.u[u];
            pu.w_off = c.conv_w[u]; pu.b_off = c.conv_b[u]; pu.g_off = c.bn_w[u]; pu.beta_off = c.bn_b[u];
            pu.N = n->uN[u]; pu.C = n->uC[u]; pu.Cp = n->uCp[u];
            pu.Wf = n->Wf[u]; pu.Wd = (training && u > 0) ? n->Wd[u] : nullptr; pu.vec = n->vec[u];
            pu.wf_gs = 9LL * n->uN[u] * n->uCp[u]; pu.wd_gs = pu.wf_gs; pu.vec_gs = 3LL * n->uN[u];
        }
        all.w_f16 = n->f16;
        int r = vv_prep_conv_w_all(n->params, n->slot, c.slot_param_stride, all, G, st);
        if (r) return r;
    } else {
        for (int u = 0; u < NU; u++) {
            int r = vv_prep_conv_w(n->params, n->slot, c.slot_param_stride, c.conv_w[u], c.conv_b[u], c.bn_w[u], c.bn_b[u], n->uN[u], n->uC[u],
                                   n->uCp[u], n->Wf[u], 9LL * n->uN[u] * n->uCp[u], (training && u > 0) ? n->Wd[u] : nullptr,
                                   9LL * n->uN[u] * n->uCp[u], n->f16, n->vec[u], 3LL * n->uN[u], G, st);
            if (r) return r;
        }
    }
    // transposed-conv weights are first needed a third of the way into the forward: re-lay them out on the side stream meanwhile
    cudaEvent_t ct_ready = nullptr;
    for (int k = 0; k < NT; k++) {
        int r = vv_prep_ct_w(n->params, n->slot, c.slot_param_stride, c.up_w[k], c.up_b[k], n->tCi[k], n->tCo[k], n->tWf[k],
                             16LL * n->tCo[k] * n->tCi[k], n->tWd[k], 16LL * n->tCo[k] * n->tCi[k], n->f16, n->tvec[k], n->tCo[k], G, sP);
        if (r) return r;
    }
    if (n->use_side) {
        ct_ready = n->ev[VV_NEV - 2];
        VV_CK(cudaEventRecord(ct_ready, sP));
    }
    if (training) VV_CK(cudaMemsetAsync(n->zero_fwd, 0, n->zero_fwd_bytes, st));
    n->bwd_zeroed = 0;
    if (training && n->use_side && n->have_zero_ev) {
        // what the backward accumulates into is zeroed NOW on the side stream (94 MB of memsets, ~22 us) instead of at the head of
        // the backward on the critical path; sP already waits for everything queued before this forward (the previous Adam included)
        VV_CK(cudaMemsetAsync(n->zero_bwd, 0, n->zero_bwd_bytes, sP));
        for (int g = 0; g < G; g++)
            VV_CK(cudaMemsetAsync(n->grads + n->slot.v[g] * c.slot_param_stride, 0, c.slot_param_stride * sizeof(float), sP));
        VV_CK(cudaEventRecord(n->ev_zero, sP));
        n->bwd_zeroed = 1;
    }
    // 2. the erased-frame inputs of every UNet
    {
        int r = vv_prep_input(x, n->X0, n->f16, G, B, n->T, S, n->cinp, c.padding, n->erase, st);
        if (r) return r;
    }
    const VvTaps t3 = taps3x3(+1), t2 = taps2x2(+1);
    auto conv_unit = [&](int u) -> int {
        const View &in = f.in[u];
        const View &y = f.y[u];
        const int H = n->uH[u], N = n->uN[u];
        VvIGemm p;
        memset(&p, 0, sizeof(p));
        p.A = (const float *)in.p; p.a_gs = in.gs; p.lda = in.ld; p.a_coff = in.coff; p.a_s2d = 0; p.Kt = n->uCp[u];
        p.B = B; p.H = H; p.W = H; p.ab_f16 = n->f16;
        p.Wt = (const float *)n->Wf[u]; p.w_gs = 9LL * N * n->uCp[u]; p.taps = t3; p.N = N;
        p.O = (float *)n->Z[u]; p.o_gs = (long long)B * H * H * N; p.ldo = N; p.o_coff = 0; p.o_d2s = 0; p.o_f16 = n->f16;
        p.bias = n->vec[u]; p.bias_gs = 3LL * N;
        p.stats = training ? n->stats[u] : nullptr; p.stats_gs = 2LL * N;
        p.G = G;
        int r = run_igemm(n->cfg.use_tensor_cores != 0, p, st, n->uC[u]);
        if (r) return r;
        VvProfScope ps(VV_PROF_BN, 0, st);
        VvBnApply q;
        memset(&q, 0, sizeof(q));
        q.Z = n->Z[u]; q.z_gs = p.o_gs;
        q.Y = y.p; q.y_gs = y.gs; q.ldy = y.ld; q.y_coff = y.coff; q.y_f16 = n->f16;
        q.pool = f.pool[u] >= 0;
        if (q.pool) { q.P = n->PL[f.pool[u]]; q.p_gs = (long long)B * (H / 2) * (H / 2) * N; }
        q.M = B * H * H; q.H = H; q.W = H; q.C = N; q.training = training;
        q.rev = n->serp;                                    // conv tiles walked the rows upwards: come back down (L2 reuse)
        q.stats = n->stats[u]; q.stats_gs = 2LL * N;
        q.vec = n->vec[u]; q.vec_gs = 3LL * N;
        q.running = n->running; q.slot = n->slot; q.slot_stat_stride = c.slot_stat_stride; q.rm_off = c.run_mean[u]; q.rv_off = c.run_var[u];
        q.save = training ? n->save[u] : nullptr; q.save_gs = 4LL * N;
        return vv_bn_apply(q, G, st);
    };
    auto convT = [&](int k) -> int {
        const View &in = f.tin[k];
        const View &o = f.tout[k];
        VvIGemm p;
        memset(&p, 0, sizeof(p));
        p.A = (const float *)in.p; p.a_gs = in.gs; p.lda = in.ld; p.a_coff = in.coff; p.a_s2d = 0; p.Kt = n->tCi[k];
        p.B = B; p.H = n->tH[k]; p.W = n->tH[k]; p.ab_f16 = n->f16;
        p.Wt = (const float *)n->tWf[k]; p.w_gs = 16LL * n->tCo[k] * n->tCi[k]; p.taps = t2; p.N = 4 * n->tCo[k];
        p.O = (float *)o.p; p.o_gs = o.gs; p.ldo = o.ld; p.o_coff = o.coff; p.o_d2s = 1; p.o_f16 = n->f16;
        p.bias = n->tvec[k]; p.bias_gs = n->tCo[k];
        p.stats = nullptr; p.G = G;
        // algorithmic FLOPs of ConvTranspose2d(k3, s2): 9 taps x Ci x Co per INPUT pixel
        return run_igemm(n->cfg.use_tensor_cores != 0, p, st, 0, igemm_flops(B, n->tH[k], n->tH[k], n->tCo[k], n->tCi[k], 9, G));
    };
    int r;
    for (int u = 0; u < 8; u++) if ((r = conv_unit(u))) return r;
    for (int k = 0; k < 3; k++) {
        if (k == 0 && ct_ready) VV_CK(cudaStreamWaitEvent(st, ct_ready, 0));
        if ((r = convT(k))) return r;
        if ((r = conv_unit(8 + 2 * k))) return r;
        if ((r = conv_unit(9 + 2 * k))) return r;
    }
    // 3. 1x1 output conv (+ squared error / MSE gradient)
    VvOutFwd q;
    memset(&q, 0, sizeof(q));
    q.U = n->UU[2]; q.u_gs = (long long)B * S * S * F; q.u_f16 = n->f16;
    q.params = n->params; q.slot = n->slot; q.slot_param_stride = c.slot_param_stride; q.w_off = c.out_w; q.b_off = c.out_b;
    q.out_channels = n->outc; q.target_is_flow = n->isflow; q.target_index = n->tidx; q.out_slot = n->oslot;
    q.B = B; q.S = S; q.F = F;
    q.raw_out = raw_out; q.raw_out_channels = raw_out_channels; q.of_out = of_out; q.of_out_channels = of_out_channels;
    q.x = x; q.x_channels = 3 * n->T; q.x_of = x_of; q.x_of_channels = x_of_channels;
    q.sse = sse;
    q.dout = (sse && training) ? n->DOUT : nullptr;
    // loss = lambda_raw * mean_{B,3*n_raw,S,S} + lambda_of * mean_{B,2*n_of,S,S}   (train.py:385-392)
    q.coef_raw = n->n_raw_tot ? 2.f * lambda_raw / ((float)B * 3.f * n->n_raw_tot * S * S) : 0.f;
    q.coef_of = n->n_of_tot ? 2.f * lambda_of / ((float)B * 2.f * n->n_of_tot * S * S) : 0.f;
    if ((r = vv_outconv_fwd(q, G, st))) return r;
    // fp16 gradient operands: one power-of-two loss scale for the whole backward, from the largest MSE-gradient coefficient
    // (|d loss / d out| = coef * |out - tgt|, residuals are O(1)): dZ values land around 2^3, 12 binades under the fp16 maximum
    // and 17 above its smallest normal.  The scale is applied once (where the last unit's dZ is stored) and removed wherever a
    // parameter gradient is written, so gradients come out unscaled.
    n->lscale = 1.f;
    if (n->f16) {
        const float cm = q.coef_raw > q.coef_of ? q.coef_raw : q.coef_of;
        n->lscale = n->lscale_user > 0.f ? n->lscale_user : (cm > 0.f ? exp2f(floorf(log2f(1.f / cm)) + 3.f) : 1.f);
    }
    n->lastB = B; n->last_training = training; n->have_dout = (sse && training) ? 1 : 0;
    return 0;
}

extern "C" int vecvad_net_losses(vecvad_net *n, const float *sse, int batch, float *losses, vecvad_stream stream) {
    VV_REQUIRE(n && sse && losses && batch >= 1, "net_losses: bad arguments");
    float inv_raw = n->n_raw_tot ? 1.f / ((float)batch * 3.f * n->n_raw_tot * n->S * n->S) : 0.f;
    float inv_of = n->n_of_tot ? 1.f / ((float)batch * 2.f * n->n_of_tot * n->S * n->S) : 0.f;
    return vv_losses(sse, n->G, batch, n->isflow, inv_raw, inv_of, losses, (cudaStream_t)stream);
}

extern "C" int vecvad_net_backward(vecvad_net *n, const float *grad_raw_out, const float *grad_of_out, vecvad_stream stream) {
    VV_REQUIRE(n && n->ws && n->grads, "net_backward: net not bound (or no gradient buffer)");
    VV_REQUIRE(n->lastB > 0 && n->last_training, "net_backward: no training-mode forward to differentiate");
    const bool ext = grad_raw_out || grad_of_out;
    VV_REQUIRE(ext || n->have_dout, "net_backward: no output gradient (pass grad_*_out or run forward with sse)");
    if (ext) {
        VV_REQUIRE(n->n_raw_out == 0 || grad_raw_out, "net_backward: grad_raw_out missing");
        VV_REQUIRE(n->n_of_out == 0 || grad_of_out, "net_backward: grad_of_out missing");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int G = n->G, B = n->lastB, F = n->F, S = n->S;
    const vecvad_net_config &c = n->cfg;
    Flow f;
    wire(n, B, f);
    int r;
    // every gradient of every slot is (re)written: zero first, then kernels accumulate / overwrite
    long long max_slot = 0;
    for (int g = 0; g < G; g++) max_slot = n->slot.v[g] > max_slot ? n->slot.v[g] : max_slot;
    (void)max_slot;
    if (n->bwd_zeroed) {                                       // zeroed on the side stream during the forward
        VV_CK(cudaStreamWaitEvent(st, n->ev_zero, 0));
        n->bwd_zeroed = 0;
    } else {
        VV_CK(cudaMemsetAsync(n->zero_bwd, 0, n->zero_bwd_bytes, st));
        for (int g = 0; g < G; g++)
            VV_CK(cudaMemsetAsync(n->grads + n->slot.v[g] * c.slot_param_stride, 0, c.slot_param_stride * sizeof(float), st));
    }

    // ---- output conv.  With the loss gradient staged by the forward (no external gradients) its backward is folded into the BatchNorm
    // backward of the last conv unit (VvBnBwd::dout): dU is never written or read.
    const bool fuse_out = !ext;
    const int f16 = n->f16;
    const float LS = f16 ? n->lscale : 1.f, inv_LS = 1.f / LS;      // loss scale carried by every dY / dZ behind the last unit (fp16 mode)
    if (!fuse_out) {
        VvOutBwd q;
        memset(&q, 0, sizeof(q));
        q.U = n->UU[2]; q.u_gs = (long long)B * S * S * F; q.u_f16 = f16;
        q.dU = n->GA; q.du_gs = q.u_gs; q.du_scale = LS;       // fp16 mode: dU enters the chain already loss-scaled
        q.params = n->params; q.grads = n->grads; q.slot = n->slot; q.slot_param_stride = c.slot_param_stride; q.w_off = c.out_w; q.b_off = c.out_b;
        q.out_channels = n->outc; q.target_is_flow = n->isflow; q.out_slot = n->oslot;
        q.M = B * S * S; q.S = S; q.F = F;
        q.dout = n->DOUT;
        q.grad_raw_out = grad_raw_out; q.raw_out_channels = 3 * n->n_raw_tot;
        q.grad_of_out = grad_of_out; q.of_out_channels = 2 * n->n_of_tot;
        if ((r = vv_outconv_bwd(q, G, st))) return r;
    }
    const VvTaps t3f = taps3x3(+1), t3b = taps3x3(-1), t2f = taps2x2(+1), t2b = taps2x2(-1);
    // ---- side stream for the weight-gradient tiles.  fork: sB waits for everything issued so far on st, runs the wgrad, and
    // leaves an event; a kernel on st that overwrites a buffer the wgrad still reads (fp32 mode: GA / GB, where dZ replaces dY in
    // place; fp16 mode: the ring of dZ operand buffers) waits for it first.
    cudaStream_t sB = n->use_side ? n->wg_stream : st;
    constexpr int NPEND = 6;
    const void *pend_ptr[NPEND] = {n->GA, n->GB, n->HZ[0], n->HZ[1], n->HZ[2], n->HZ[3]};
    cudaEvent_t pend[NPEND] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    n->ev_next = 0;
    auto next_ev = [&]() { return n->ev[(n->ev_next++) % VV_NEV]; };
    auto pend_slot = [&](const void *ptr) -> int {
        for (int i = 0; i < NPEND; i++)
            if (ptr && ptr == pend_ptr[i]) return i;
        return -1;
    };
    auto before_write = [&](const void *ptr) -> int {
        const int i = pend_slot(ptr);
        if (n->use_side && i >= 0 && pend[i]) {
            VV_CK(cudaStreamWaitEvent(st, pend[i], 0));
            pend[i] = nullptr;
        }
        return 0;
    };
    auto fork = [&]() -> int {
        if (!n->use_side) return 0;
        cudaEvent_t e = next_ev();
        VV_CK(cudaEventRecord(e, st));
        VV_CK(cudaStreamWaitEvent(sB, e, 0));
        return 0;
    };
    auto mark = [&](const void *reads) -> int {
        if (!n->use_side) return 0;
        cudaEvent_t e = next_ev();
        VV_CK(cudaEventRecord(e, sB));
        const int i = pend_slot(reads);
        if (i >= 0) pend[i] = e;
        return 0;
    };
    int hz_next = 0;
    bool batched_scatter = true;       // conv weight gradients go back to PyTorch's layout in one launch at the end
    for (int u = 0; u < NU; u++) batched_scatter = batched_scatter && n->uN[u] % 32 == 0 && n->uCp[u] % 32 == 0;

    // gradient of the post-ReLU output of unit u lives in dyv (fp32); produces d(input of unit u) into `din` (if u > 0).
    // dz_inplace: where dZ goes in the fp32 modes (it may overwrite dY); the fp16 mode takes the next buffer of its dZ ring.
    // up_k >= 0: unit u is the first conv of up-block up_k, whose input gradient splits into the skip half (din, fp32) and the
    // transposed conv's output gradient (fp16 mode: stored as fp16 in dUP[up_k], the operand of the two transposed-conv backward tiles).
    auto unit_bwd = [&](int u, const View &dyv, void *dz_inplace, const View *din, int up_k) -> int {
        const int H = n->uH[u], N = n->uN[u], M = B * H * H;
        void *dz_buf = f16 ? n->HZ[(hz_next++) & 3] : dz_inplace;
        VvBnBwd q;
        memset(&q, 0, sizeof(q));
        q.Z = n->Z[u]; q.z_gs = (long long)M * N; q.z_f16 = f16;
        q.dY = dyv.p; q.dy_gs = dyv.gs; q.ldy = dyv.ld; q.dy_coff = dyv.coff;
        q.dZ = dz_buf; q.dz_gs = (long long)M * N; q.dz_f16 = f16;
        // serpentine row order: every kernel of the backward chain walks against its predecessor, so what was written / read last
        // (still in L2) is read first: dgrad of unit u+1 -> reduce(u) -> apply(u) -> dgrad(u) alternate, hence the parity of u
        const int up = n->serp ? (u & 1) : 0;
        q.rev_reduce = n->serp ? !up : 0; q.rev_apply = up;
        const bool unscaled_dy = fuse_out && u == NU - 1;      // the fused last unit forms dY from the (unscaled) staged loss gradient:
        q.store_scale = unscaled_dy ? LS : 1.f;                // the loss scale enters where its dZ is stored ...
        q.grad_unscale = unscaled_dy ? 1.f : inv_LS;           // ... every other dY carries it already: removed from d gamma / d beta
        q.M = M; q.C = N;
        q.save = n->save[u]; q.save_gs = 4LL * N;
        q.sums = n->bsums[u]; q.sums_gs = 2LL * N;
        q.grads = n->grads; q.slot = n->slot; q.slot_param_stride = c.slot_param_stride; q.gamma_off = c.bn_w[u]; q.beta_off = c.bn_b[u];
        if (fuse_out && u == NU - 1) {
            q.dout = n->DOUT; q.params = n->params; q.ow_off = c.out_w; q.ob_off = c.out_b; q.out_channels = n->outc;
        }
        int rr;
        if ((rr = before_write(dz_buf))) return rr;
        {
            VvProfScope ps(VV_PROF_BN, 0, st);
            rr = vv_bn_bwd(q, G, st);
        }
        if (rr) return rr;
        // weight gradient
        const View &in = f.in[u];
        VvWGrad w;
        memset(&w, 0, sizeof(w));
        w.A = (const float *)in.p; w.a_gs = in.gs; w.lda = in.ld; w.a_coff = in.coff; w.Kt = n->uCp[u];
        w.B = B; w.H = H; w.W = H; w.ab_f16 = f16;
        w.Gd = (const float *)dz_buf; w.g_gs = (long long)M * N; w.ldg = N; w.g_coff = 0; w.g_s2d = 0; w.N = N;
        w.taps = t3f; w.dW = n->dWf[u]; w.dw_gs = 9LL * N * n->uCp[u]; w.G = G;
        if ((rr = fork())) return rr;
        if ((rr = run_wgrad(c.use_tensor_cores != 0, w, sB, n->uC[u]))) return rr;
        if (!batched_scatter &&
            (rr = vv_scatter_conv_wgrad(n->dWf[u], w.dw_gs, N, n->uC[u], n->uCp[u], n->grads, n->slot, c.slot_param_stride, c.conv_w[u], G, sB)))
            return rr;
        if ((rr = mark(dz_buf))) return rr;
        // (pre-BN conv bias: its gradient is exactly zero in training mode -- left at the memset value; the reference
        //  produces round-off noise there)
        if (din) {
            VvIGemm p;
            memset(&p, 0, sizeof(p));
            p.A = (const float *)dz_buf; p.a_gs = (long long)M * N; p.lda = N; p.a_coff = 0; p.a_s2d = 0; p.Kt = N; p.ab_f16 = f16;
            p.B = B; p.H = H; p.W = H;
            p.Wt = (const float *)n->Wd[u]; p.w_gs = 9LL * N * n->uCp[u]; p.taps = t3b; p.N = n->uC[u];
            p.O = (float *)din->p; p.o_gs = din->gs; p.ldo = din->ld; p.o_coff = din->coff; p.o_d2s = 0; p.o_f16 = f16;
            if (f16 && up_k >= 0) {
                p.o_split = n->uC[u] / 2; p.O2 = n->dUP[up_k]; p.ldo2 = n->uC[u] / 2; p.o2_gs = (long long)M * (n->uC[u] / 2);
            }
            p.bias = nullptr; p.stats = nullptr; p.G = G;
            p.rev = n->serp ? !up : 0;                        // apply(u) walked up (up = 1) or down: the input-gradient tiles go the other way
            if ((rr = before_write(din->p))) return rr;
            if ((rr = run_igemm(c.use_tensor_cores != 0, p, st))) return rr;
        }
        return 0;
    };
    // transposed conv k: its output gradient is the second half of the concat gradient: fp32 modes: columns [Co, 2Co) of dCAT[k];
    // fp16 mode: the dense fp16 tensor dUP[k] the split epilogue of the producing tile wrote
    auto convT_bwd = [&](int k, const View &ddeep) -> int {
        const int Hi = n->tH[k], Ci = n->tCi[k], Co = n->tCo[k];
        const int Hc = 2 * Hi, Cc = 2 * Co;                   // concat buffer geometry
        View dhalf = f16 ? mk(n->dUP[k], B, Hc, Co, 0, Co) : mk(n->dCAT[k], B, Hc, Cc, Co, Co);
        int rr = vv_colsum(dhalf.p, f16, dhalf.gs, dhalf.ld, dhalf.coff, B * Hc * Hc, Co, inv_LS, n->grads, n->slot, c.slot_param_stride, c.up_b[k], G, st);
        if (rr) return rr;
        const View &in = f.tin[k];
        VvWGrad w;
        memset(&w, 0, sizeof(w));
        w.A = (const float *)in.p; w.a_gs = in.gs; w.lda = in.ld; w.a_coff = in.coff; w.Kt = Ci;
        w.B = B; w.H = Hi; w.W = Hi; w.ab_f16 = f16;
        w.Gd = (const float *)dhalf.p; w.g_gs = dhalf.gs; w.ldg = dhalf.ld; w.g_coff = dhalf.coff; w.g_s2d = 1; w.N = 4 * Co;
        w.taps = t2f; w.dW = n->tdW[k]; w.dw_gs = 16LL * Co * Ci; w.G = G;
        if ((rr = fork())) return rr;
        if ((rr = run_wgrad(c.use_tensor_cores != 0, w, sB, 0, igemm_flops(B, Hi, Hi, Co, Ci, 9, G)))) return rr;
        if ((rr = vv_scatter_ct_wgrad(n->tdW[k], w.dw_gs, Ci, Co, inv_LS, n->grads, n->slot, c.slot_param_stride, c.up_w[k], G, sB))) return rr;
        if ((rr = mark(nullptr))) return rr;
        VvIGemm p;
        memset(&p, 0, sizeof(p));
        p.A = (const float *)dhalf.p; p.a_gs = dhalf.gs; p.lda = dhalf.ld; p.a_coff = dhalf.coff; p.a_s2d = 1; p.Kt = 4 * Co; p.ab_f16 = f16;
        p.B = B; p.H = Hi; p.W = Hi;
        p.Wt = (const float *)n->tWd[k]; p.w_gs = 16LL * Co * Ci; p.taps = t2b; p.N = Ci;
        p.O = (float *)ddeep.p; p.o_gs = ddeep.gs; p.ldo = ddeep.ld; p.o_coff = ddeep.coff; p.o_d2s = 0; p.o_f16 = f16;
        p.bias = nullptr; p.stats = nullptr; p.G = G;
        if ((rr = before_write(ddeep.p))) return rr;
        return run_igemm(c.use_tensor_cores != 0, p, st, 0, igemm_flops(B, Hi, Hi, Co, Ci, 9, G));
    };

    // conv weight gradients of units [u0, u1) back to PyTorch's layout (one launch) on stream s
    auto scatter_units = [&](int u0, int u1, cudaStream_t s) -> int {
        VvPrepAll all;
        memset(&all, 0, sizeof(all));
        all.n = u1 - u0;
        all.scale = inv_LS;
        for (int u = u0; u < u1; u++) {
            VvPrepUnit &pu = all.u[u - u0];
            pu.w_off = c.conv_w[u]; pu.N = n->uN[u]; pu.C = n->uC[u]; pu.Cp = n->uCp[u];
            pu.dWf = n->dWf[u]; pu.wf_gs = 9LL * n->uN[u] * n->uCp[u];
        }
        return vv_scatter_conv_wgrad_all(n->grads, n->slot, c.slot_param_stride, all, G, s);
    };
    // gradient phase ph is final once everything issued so far on BOTH streams has run: the side stream (which carries the weight
    // gradients of the phase and their scatter) waits for the main stream's share (BatchNorm / bias / output-conv gradients)
    // and records the phase event.  Nothing on the main stream waits for it.
    n->phases_recorded = 0;
    auto phase_done = [&](int ph, int u0, int u1) -> int {
        int rr;
        if (batched_scatter && (rr = scatter_units(u0, u1, sB))) return rr;
        if (!n->have_phase_ev) return 0;
        if (n->use_side) {
            cudaEvent_t e = next_ev();
            VV_CK(cudaEventRecord(e, st));
            VV_CK(cudaStreamWaitEvent(sB, e, 0));
        }
        VV_CK(cudaEventRecord(n->ev_phase[ph], sB));
        n->phases_recorded = ph + 1;
        return 0;
    };

    // ---- decoder, deepest last.  GA holds dU3 now (external gradients), or nothing (fused: formed from DOUT on the fly).
    void *ga = n->GA, *gb = n->GB;
    for (int k = 2; k >= 0; k--) {
        const int H = S >> (2 - k), C = F << (3 - k);          // concat geometry of up-block k
        const int u2 = 9 + 2 * k, u1 = 8 + 2 * k;
        View dy2 = mk(ga, B, H, C / 2, 0, C / 2);               // d(output of second conv)
        View dmid = mk(gb, B, H, C / 2, 0, C / 2);
        if ((r = unit_bwd(u2, dy2, ga, &dmid, -1))) return r;    // fp32: dz in place in ga; d(mid) -> gb
        View dcat = mk(n->dCAT[k], B, H, C, 0, C);
        if ((r = unit_bwd(u1, dmid, gb, &dcat, k))) return r;    // fp32: dz in place in gb; d(concat) -> dCAT[k] (+ dUP[k])
        View ddeep = mk(ga, B, H / 2, C, 0, C);                  // d(X4 / U1 / U2): [B,(H/2)^2, C]
        if ((r = convT_bwd(k, ddeep))) return r;
    }
    if ((r = phase_done(0, 8, NU))) return r;                  // decoder (units 8..13, transposed convs, output conv) complete
    // ---- encoder.  ga holds dX4.
    for (int k = 3; k >= 0; k--) {
        if (k == 2 && (r = phase_done(1, 6, 8))) return r;     // deepest encoder block (units 6, 7) complete
        const int u2 = 2 * k + 1, u1 = 2 * k;
        const int H = n->uH[u1], N = n->uN[u2];
        View dy2;
        if (k == 3) dy2 = mk(ga, B, H, N, 0, N);
        else dy2 = mk(n->dCAT[2 - k], B, H, 2 * N, 0, N);       // first half of the concat gradient (+ pooled path, added below)
        // fp32 modes: dZ of the second conv replaces its dY in place (ga), or goes to gb when dY sits in dCAT; d(mid) takes the other
        // scratch buffer and d(pooled input) the first again.  fp16 mode: dZ lives in the ring, so d(mid) -> gb, d(pooled input) -> ga.
        void *dz2 = (k == 3) ? ga : gb;
        void *other = f16 ? gb : ((dz2 == ga) ? gb : ga);
        void *pool_buf = f16 ? ga : dz2;
        View dmid = mk(other, B, H, N, 0, N);
        if ((r = unit_bwd(u2, dy2, dz2, &dmid, -1))) return r;
        if (k == 0) {
            if ((r = unit_bwd(u1, dmid, other, nullptr, -1))) return r;   // first conv: its input needs no gradient
        } else {
            // d(pooled input) -> scratch, then routed through the max-pool into the skip half of dCAT[3-k]
            View dpool = mk(pool_buf, B, H, n->uC[u1], 0, n->uC[u1]);
            if ((r = unit_bwd(u1, dmid, other, &dpool, -1))) return r;
            const int Hs = 2 * H, Cs = n->uC[u1];                 // skip tensor geometry (x_k) : [B,Hs,Hs,Cs] inside CAT[3-k]
            View ysk = mk(n->CAT[3 - k], B, Hs, 2 * Cs, 0, Cs);
            View dsk = mk(n->dCAT[3 - k], B, Hs, 2 * Cs, 0, Cs);
            if ((r = vv_maxpool_bwd(ysk.p, f16, ysk.gs, ysk.ld, ysk.coff, dpool.p, dpool.gs, dsk.p, dsk.gs, dsk.ld, dsk.coff, G, B,
                                    Hs, Hs, Cs, st)))
                return r;
        }
    }
    if ((r = phase_done(2, 0, 6))) return r;                   // the rest; its scatter runs on the side stream like the others
    if (n->use_side && !(n->defer_join && n->have_phase_ev)) { // join: every weight gradient and scatter is complete
        if (n->have_phase_ev) VV_CK(cudaStreamWaitEvent(st, n->ev_phase[2], 0));
        else {
            cudaEvent_t e = next_ev();
            VV_CK(cudaEventRecord(e, sB));
            VV_CK(cudaStreamWaitEvent(st, e, 0));
        }
    }
    return 0;
}

extern "C" int vecvad_net_defer_join(vecvad_net *n, int defer) {
    VV_REQUIRE(n, "defer_join: null net");
    n->defer_join = defer ? 1 : 0;
    return 0;
}

extern "C" int vecvad_net_grad_phase_ranges(const vecvad_net *n, int64_t *begin, int64_t *end) {
    VV_REQUIRE(n && begin && end, "grad_phase_ranges: null argument");
    const vecvad_net_config &c = n->cfg;
    begin[0] = c.conv_w[8]; end[0] = c.slot_param_stride;
    begin[1] = c.conv_w[6]; end[1] = c.conv_w[8];
    begin[2] = 0;           end[2] = c.conv_w[6];
    // the phases rely on the slot layout of vec_vad_b200/unet.py slot_layout(): units in execution order, then the transposed
    // convs, then the output conv
    for (int u = 1; u < NU; u++) VV_REQUIRE(c.conv_w[u] > c.conv_w[u - 1], "grad_phase_ranges: conv units are not laid out in execution order");
    VV_REQUIRE(c.up_w[0] > c.conv_w[NU - 1] && c.out_w > c.up_w[0], "grad_phase_ranges: transposed / output convs do not follow the conv units");
    return 0;
}

extern "C" int vecvad_net_grad_phase_wait(vecvad_net *n, int phase, vecvad_stream stream) {
    VV_REQUIRE(n && phase >= 0 && phase < 3, "grad_phase_wait: phase %d out of range", phase);
    VV_REQUIRE(n->have_phase_ev && n->phases_recorded > phase, "grad_phase_wait: phase %d was not recorded by the last backward", phase);
    VV_CK(cudaStreamWaitEvent((cudaStream_t)stream, n->ev_phase[phase], 0));
    return 0;
}

extern "C" int vecvad_net_set_loss_scale(vecvad_net *n, float scale) {
    VV_REQUIRE(n, "set_loss_scale: null net");
    if (scale > 0.f) {
        int e;
        VV_REQUIRE(frexpf(scale, &e) == 0.5f, "set_loss_scale: %g is not a power of two", (double)scale);
    }
    n->lscale_user = scale;
    return 0;
}

extern "C" int vecvad_net_debug_read(vecvad_net *n, int kind, int index, float *dst, int64_t max_floats, int64_t *n_floats,
                                     vecvad_stream stream) {
    VV_REQUIRE(n && n->ws && dst && n_floats && n->lastB > 0, "debug_read: net not bound / no forward yet");
    const long long G = n->G, B = n->lastB, F = n->F, S = n->S;
    auto M = [&](long long H) { return B * H * H; };
    const void *src = nullptr;
    long long cnt = 0;
    bool half = false;                       // the buffer holds fp16 in the fp16-operand mode: converted on the way out
    const int u = index, k = index;
    switch (kind) {
        case 0: src = n->X0; cnt = G * M(S) * n->cinp; half = true; break;
        case 1: VV_REQUIRE(u >= 0 && u < NU, "debug_read: unit"); src = n->Z[u]; cnt = G * M(n->uH[u]) * n->uN[u]; half = true; break;
        case 2: VV_REQUIRE(u >= 0 && u < NU && u % 2 == 0, "debug_read: unit"); src = n->A[u]; cnt = G * M(n->uH[u]) * n->uN[u]; half = true; break;
        case 3: VV_REQUIRE(k >= 0 && k < 3, "debug_read: k"); src = n->CAT[k]; cnt = G * M(S >> (2 - k)) * (F << (3 - k)); half = true; break;
        case 4: VV_REQUIRE(k >= 0 && k < 3, "debug_read: k"); src = n->PL[k]; cnt = G * M(S >> (k + 1)) * (F << k); half = true; break;
        case 5: src = n->X4; cnt = G * M(S >> 3) * 8 * F; half = true; break;
        case 6: VV_REQUIRE(k >= 0 && k < 3, "debug_read: k"); src = n->UU[k]; cnt = G * M(S >> (2 - k)) * (F << (2 - k)); half = true; break;
        case 7: VV_REQUIRE(k >= 0 && k < 3, "debug_read: k"); src = n->dCAT[k]; cnt = G * M(S >> (2 - k)) * (F << (3 - k)); half = true; break;
        case 8: src = n->GA; cnt = G * M(S) * F; half = true; break;
        case 9: src = n->GB; cnt = G * M(S) * F; half = true; break;
        case 10: src = n->DOUT; cnt = G * M(S) * 4; break;
        case 11: VV_REQUIRE(u >= 0 && u < NU, "debug_read: unit"); src = n->Wf[u]; cnt = G * 9 * n->uN[u] * n->uCp[u]; half = true; break;
        case 12: VV_REQUIRE(u >= 0 && u < NU, "debug_read: unit"); src = n->dWf[u]; cnt = G * 9 * n->uN[u] * n->uCp[u]; break;
        case 13: VV_REQUIRE(k >= 0 && k < NT, "debug_read: k"); src = n->tWf[k]; cnt = G * 16 * n->tCo[k] * n->tCi[k]; half = true; break;
        case 14: VV_REQUIRE(k >= 0 && k < NT, "debug_read: k"); src = n->tdW[k]; cnt = G * 16 * n->tCo[k] * n->tCi[k]; break;
        case 15: VV_REQUIRE(k >= 0 && k < 3 && n->f16, "debug_read: dUP exists in the fp16 mode only"); src = n->dUP[k];
                 cnt = G * M(S >> (2 - k)) * (F << (2 - k)); half = true; break;
        case 16: VV_REQUIRE(k >= 0 && k < 4 && n->f16, "debug_read: the dZ ring exists in the fp16 mode only"); src = n->HZ[k]; cnt = G * M(S) * F; half = true; break;
        default: return vv_set_err(-1, "debug_read: unknown kind %d", kind);
    }
    VV_REQUIRE(cnt <= max_floats, "debug_read: destination too small (%lld > %lld)", cnt, (long long)max_floats);
    *n_floats = cnt;
    if (half && n->f16) return vv_f16_to_f32(src, cnt, dst, (cudaStream_t)stream);
    VV_CK(cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
