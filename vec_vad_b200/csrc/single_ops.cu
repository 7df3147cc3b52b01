// Single contractions on NHWC tensors through the C ABI (include/vecvad.h, "single ops"): exported for unit tests and for profiling
// one kernel at a time.  They run exactly the launches the net engine issues for the same shape (vv_run_igemm / vv_run_wgrad in net.cu).
//
// use_tc (all entry points):  0 = fp32 SIMT tiles; low nibble 1 = the tcgen05 tile the engine picks for the shape, 3 = flattened-
// sequence tiles, 4 (forward) / 2 (wgrad) = pair / tap-reuse tiles; + 16 = fp16 operands (kind::f16, fp32 accumulation): operands are
// converted to fp16 inside the call into stream-ordered temporaries (cudaMallocAsync; these are test entry points, the net engine
// itself never allocates), outputs stay fp32.
#include "unet_kernels.h"

int vv_run_igemm(bool want_tc, const VvIGemm &p, cudaStream_t st);
int vv_run_wgrad(bool want_tc, const VvWGrad &p, cudaStream_t st);
VvTaps vv_taps3x3(int sign);
VvTaps vv_taps2x2(int sign);

namespace {

struct Tmp {                     // stream-ordered scratch, released when the entry point returns (after its launches are queued)
    cudaStream_t st;
    void *ptr[6];
    int n;
    explicit Tmp(cudaStream_t s) : st(s), n(0) {}
    void *get(long long bytes) {
        void *p = nullptr;
        if (n >= 6 || cudaMallocAsync(&p, (size_t)(bytes + 256), st) != cudaSuccess) return nullptr;
        ptr[n++] = p;
        return p;
    }
    ~Tmp() {
        for (int i = 0; i < n; i++) cudaFreeAsync(ptr[i], st);
    }
};

// fp32 [rows][ld] (channels coff .. coff+cols) -> dense fp16 [rows][cols]
const float *to_half(Tmp &t, const float *src, int ld, int coff, int cols, long long rows, int *err) {
    void *h = t.get(rows * cols * 2);
    if (!h) { *err = vv_set_err(-2, "single op: cudaMallocAsync failed"); return nullptr; }
    *err = vv_f32_to_f16(src + coff, ld, cols, rows, h, t.st);
    return (const float *)h;
}

int pick_igemm(int use_tc, const VvIGemm &p, cudaStream_t st, const char *what) {
    const int base = use_tc & 15;
    if (base == 0) return vv_launch_igemm_simt(p, st);
    VV_REQUIRE(vv_igemm_tc_supported(p), "%s: shape not supported by the tcgen05 path", what);
    if (base == 3) {
        VV_REQUIRE(vv_igemm_flat_shape_ok(p), "%s: shape not supported by the flattened-sequence tiles", what);
        return vv_launch_igemm_flat(p, st);
    }
    if (base == 4) {
        VV_REQUIRE(vv_igemm_tc3_supported(p), "%s: shape not supported by the pair tiles", what);
        return vv_launch_igemm_tc3(p, st);
    }
    return vv_run_igemm(true, p, st);
}

int pick_wgrad(int use_tc, const VvWGrad &w, cudaStream_t st, const char *what) {
    const int base = use_tc & 15;
    if (base == 0) return vv_launch_wgrad_simt(w, st);
    VV_REQUIRE(vv_wgrad_tc_supported(w), "%s: shape not supported by the tcgen05 path", what);
    if (base == 3) {
        VV_REQUIRE(vv_wgrad_flat_shape_ok(w), "%s: shape not supported by the flattened-sequence tiles", what);
        return vv_launch_wgrad_flat(w, st);
    }
    if (base == 2) {
        VV_REQUIRE(vv_wgrad_tc2_supported(w), "%s: shape not supported by the tap-reuse tiles", what);
        return vv_launch_wgrad_tc2(w, st);
    }
    return vv_run_wgrad(true, w, st);
}

}  // namespace

// 3x3 pad-1 convolution forward: scratch >= 9*cout*cin floats
extern "C" int vecvad_conv3x3_forward(const float *in, int ld_in, const float *w, const float *bias, float *out, double *stats,
                                      float *scratch, int batch, int h, int wd, int cin, int cout, int use_tc, vecvad_stream stream) {
    VV_REQUIRE(in && w && out && scratch, "conv3x3_forward: null argument");
    VV_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && ld_in >= cin && ld_in % 4 == 0, "conv3x3_forward: cin/cout must be multiples of 16");
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = (use_tc & 16) != 0;
    VV_REQUIRE(!f16 || (use_tc & 15), "conv3x3_forward: fp16 operands need a tcgen05 mode");
    Tmp tmp(st);
    VvIntG slot;
    memset(&slot, 0, sizeof(slot));
    int r = vv_prep_conv_w(w, slot, 0, 0, 0, 0, 0, cout, cin, cin, scratch, 0, nullptr, 0, f16, nullptr, 0, 1, st);
    if (r) return r;
    if (stats) VV_CK(cudaMemsetAsync(stats, 0, 2 * cout * sizeof(double), st));
    VvIGemm p;
    memset(&p, 0, sizeof(p));
    p.A = in; p.a_gs = 0; p.lda = ld_in; p.Kt = cin; p.B = batch; p.H = h; p.W = wd;
    p.Wt = scratch; p.taps = vv_taps3x3(+1); p.N = cout;
    p.O = out; p.ldo = cout; p.bias = bias; p.stats = stats; p.G = 1;
    if (f16) {
        p.A = to_half(tmp, in, ld_in, 0, cin, (long long)batch * h * wd, &r);
        if (r) return r;
        p.lda = cin; p.ab_f16 = 1;
    }
    return pick_igemm(use_tc, p, st, "conv3x3_forward");
}

// weight gradient of the 3x3 pad-1 convolution, dw[cout][cin][3][3] = sum_pixels grad_out x shifted in: scratch >= 9*cout*cin floats
extern "C" int vecvad_conv3x3_wgrad(const float *in, int ld_in, const float *grad_out, float *dw, float *scratch, int batch, int h, int wd,
                                    int cin, int cout, int use_tc, vecvad_stream stream) {
    VV_REQUIRE(in && grad_out && dw && scratch, "conv3x3_wgrad: null argument");
    VV_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && ld_in >= cin && ld_in % 4 == 0, "conv3x3_wgrad: cin/cout must be multiples of 16");
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = (use_tc & 16) != 0;
    VV_REQUIRE(!f16 || (use_tc & 15), "conv3x3_wgrad: fp16 operands need a tcgen05 mode");
    Tmp tmp(st);
    VV_CK(cudaMemsetAsync(scratch, 0, 9LL * cout * cin * sizeof(float), st));
    VvWGrad w;
    memset(&w, 0, sizeof(w));
    w.A = in; w.lda = ld_in; w.Kt = cin; w.B = batch; w.H = h; w.W = wd;
    w.Gd = grad_out; w.ldg = cout; w.N = cout; w.taps = vv_taps3x3(+1); w.dW = scratch; w.G = 1;
    int r = 0;
    if (f16) {
        const long long rows = (long long)batch * h * wd;
        w.A = to_half(tmp, in, ld_in, 0, cin, rows, &r);
        if (r) return r;
        w.Gd = to_half(tmp, grad_out, cout, 0, cout, rows, &r);
        if (r) return r;
        w.lda = cin; w.ab_f16 = 1;
    }
    if ((r = pick_wgrad(use_tc, w, st, "conv3x3_wgrad"))) return r;
    VvIntG slot;
    memset(&slot, 0, sizeof(slot));
    return vv_scatter_conv_wgrad(scratch, 0, cout, cin, cin, dw, slot, 0, 0, 1, st);
}

// input gradient of the 3x3 pad-1 convolution (the engine's dgrad launch: flipped taps, Wd operand): scratch >= 18*cout*cin floats
extern "C" int vecvad_conv3x3_dgrad(const float *grad_out, const float *w, float *grad_in, float *scratch, int batch, int h, int wd, int cin,
                                    int cout, int use_tc, vecvad_stream stream) {
    VV_REQUIRE(grad_out && w && grad_in && scratch, "conv3x3_dgrad: null argument");
    VV_REQUIRE(cin % 16 == 0 && cout % 16 == 0, "conv3x3_dgrad: cin/cout must be multiples of 16");
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = (use_tc & 16) != 0;
    VV_REQUIRE(!f16 || (use_tc & 15), "conv3x3_dgrad: fp16 operands need a tcgen05 mode");
    Tmp tmp(st);
    VvIntG slot;
    memset(&slot, 0, sizeof(slot));
    float *Wf = scratch, *Wd = scratch + 9LL * cout * cin;
    int r = vv_prep_conv_w(w, slot, 0, 0, 0, 0, 0, cout, cin, cin, Wf, 0, Wd, 0, f16, nullptr, 0, 1, st);
    if (r) return r;
    VvIGemm p;
    memset(&p, 0, sizeof(p));
    p.A = grad_out; p.lda = cout; p.Kt = cout; p.B = batch; p.H = h; p.W = wd;
    p.Wt = Wd; p.taps = vv_taps3x3(-1); p.N = cin;
    p.O = grad_in; p.ldo = cin; p.G = 1;
    if (f16) {
        p.A = to_half(tmp, grad_out, cout, 0, cout, (long long)batch * h * wd, &r);
        if (r) return r;
        p.ab_f16 = 1;
    }
    return pick_igemm(use_tc, p, st, "conv3x3_dgrad");
}

// ---- ConvTranspose2d(k3, s2, p1, output_padding 1) as the engine runs it (model/unet.py:54): four sub-pixel phases = a 2x2-tap conv
// over N = 4*Co columns whose epilogue pixel-shuffles into a [B,2H,2W,ld_out] tensor at channel out_coff.
// scratch: 32*Co*Ci + Co floats (forward / dgrad) or 48*Co*Ci + Co (wgrad).
static int ct_prep(const float *w, const float *bias, int ci, int co, int f16, float *scratch, cudaStream_t st, float **Wf, float **Wd, float **vec) {
    VvIntG slot;
    memset(&slot, 0, sizeof(slot));
    *Wf = scratch; *Wd = scratch + 16LL * co * ci; *vec = scratch + 32LL * co * ci;
    // vv_prep_ct_w reads weight and bias at offsets from one base pointer
    return vv_prep_ct_w(w, slot, 0, 0, bias ? (long long)(bias - w) : 0, ci, co, *Wf, 0, *Wd, 0, f16, *vec, 0, 1, st);
}

extern "C" int vecvad_convt3x3s2_forward(const float *in, const float *w, const float *bias, float *out, int ld_out, int out_coff,
                                         float *scratch, int batch, int h, int wd, int ci, int co, int use_tc, vecvad_stream stream) {
    VV_REQUIRE(in && w && bias && out && scratch, "convt_forward: null argument");
    VV_REQUIRE(ci % 32 == 0 && co % 16 == 0 && ld_out >= out_coff + co, "convt_forward: bad channel counts");
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = (use_tc & 16) != 0;
    VV_REQUIRE(!f16 || (use_tc & 15), "convt_forward: fp16 operands need a tcgen05 mode");
    Tmp tmp(st);
    float *Wf, *Wd, *vec;
    int r = ct_prep(w, bias, ci, co, f16, scratch, st, &Wf, &Wd, &vec);
    if (r) return r;
    VvIGemm p;
    memset(&p, 0, sizeof(p));
    p.A = in; p.lda = ci; p.Kt = ci; p.B = batch; p.H = h; p.W = wd;
    p.Wt = Wf; p.taps = vv_taps2x2(+1); p.N = 4 * co;
    p.O = out; p.ldo = ld_out; p.o_coff = out_coff; p.o_d2s = 1;
    p.bias = vec; p.G = 1;
    if (f16) {
        p.A = to_half(tmp, in, ci, 0, ci, (long long)batch * h * wd, &r);
        if (r) return r;
        p.ab_f16 = 1;
    }
    return pick_igemm(use_tc, p, st, "convt_forward");
}

extern "C" int vecvad_convt3x3s2_dgrad(const float *grad_out, int ld, int coff, const float *w, float *grad_in, float *scratch, int batch,
                                       int h, int wd, int ci, int co, int use_tc, vecvad_stream stream) {
    VV_REQUIRE(grad_out && w && grad_in && scratch, "convt_dgrad: null argument");
    VV_REQUIRE(ci % 32 == 0 && co % 32 == 0 && ld >= coff + co, "convt_dgrad: bad channel counts");
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = (use_tc & 16) != 0;
    VV_REQUIRE(!f16 || (use_tc & 15), "convt_dgrad: fp16 operands need a tcgen05 mode");
    Tmp tmp(st);
    float *Wf, *Wd, *vec;
    int r = ct_prep(w, nullptr, ci, co, f16, scratch, st, &Wf, &Wd, &vec);
    if (r) return r;
    VvIGemm p;
    memset(&p, 0, sizeof(p));
    p.A = grad_out; p.lda = ld; p.a_coff = coff; p.a_s2d = 1; p.Kt = 4 * co; p.B = batch; p.H = h; p.W = wd;
    p.Wt = Wd; p.taps = vv_taps2x2(-1); p.N = ci;
    p.O = grad_in; p.ldo = ci; p.G = 1;
    if (f16) {          // the engine's fp16 mode keeps this gradient as a dense fp16 [B,2H,2W,co] tensor (net.cu: dUP)
        p.A = to_half(tmp, grad_out, ld, coff, co, (long long)batch * 4 * h * wd, &r);
        if (r) return r;
        p.lda = co; p.a_coff = 0; p.ab_f16 = 1;
    }
    return pick_igemm(use_tc, p, st, "convt_dgrad");
}

extern "C" int vecvad_convt3x3s2_wgrad(const float *in, const float *grad_out, int ld, int coff, float *dw, float *scratch, int batch, int h,
                                       int wd, int ci, int co, int use_tc, vecvad_stream stream) {
    VV_REQUIRE(in && grad_out && dw && scratch, "convt_wgrad: null argument");
    VV_REQUIRE(ci % 32 == 0 && co % 32 == 0 && ld >= coff + co, "convt_wgrad: bad channel counts");
    cudaStream_t st = (cudaStream_t)stream;
    const int f16 = (use_tc & 16) != 0;
    VV_REQUIRE(!f16 || (use_tc & 15), "convt_wgrad: fp16 operands need a tcgen05 mode");
    Tmp tmp(st);
    VV_CK(cudaMemsetAsync(scratch, 0, 16LL * co * ci * sizeof(float), st));
    VvWGrad w;
    memset(&w, 0, sizeof(w));
    w.A = in; w.lda = ci; w.Kt = ci; w.B = batch; w.H = h; w.W = wd;
    w.Gd = grad_out; w.ldg = ld; w.g_coff = coff; w.g_s2d = 1; w.N = 4 * co;
    w.taps = vv_taps2x2(+1); w.dW = scratch; w.G = 1;
    int r = 0;
    if (f16) {
        w.A = to_half(tmp, in, ci, 0, ci, (long long)batch * h * wd, &r);
        if (r) return r;
        w.Gd = to_half(tmp, grad_out, ld, coff, co, (long long)batch * 4 * h * wd, &r);
        if (r) return r;
        w.ldg = co; w.g_coff = 0; w.ab_f16 = 1;
    }
    if ((r = pick_wgrad(use_tc, w, st, "convt_wgrad"))) return r;
    VvIntG slot;
    memset(&slot, 0, sizeof(slot));
    return vv_scatter_ct_wgrad(scratch, 0, ci, co, 1.f, dw, slot, 0, 0, 1, st);
}
