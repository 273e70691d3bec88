// Host side of libmetasolver_b200.so: the C ABI (include/metasolver_b200.h) and the orchestration
// of one ODE-block integration (forward) and its discretize-then-optimize gradient (backward).
//
// The step x stage loop of the reference (`RKParametricSolver.integrate`, rk_parametric.py:104-112
// and `_make_step`, rk_parametric_order{2stage2,3stage3,4stage4}.py / euler.py) becomes a fixed
// sequence of 2*stages*n_steps convolution launches on the caller's stream; every elementwise
// operation of the reference (activations, k*w*dt, y + dt*sum b_i k_i, the adjoint combinations)
// is folded into the epilogue of the convolution that produces its operand.  No host sync.
#include <cstdarg>
#include <atomic>
#include <vector>

#include "metasolver_b200.h"
#include "msb_internal.h"
#include "msb_host.h"

namespace msb {

static_assert(kMaxSlices == MSB_MAX_SOLVERS, "EpiParams slice table must match the ABI");
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -1;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- tuning options (msb_set_option / environment MSB_<NAME>) ----
static std::atomic<int> g_tune[TUNE_COUNT];
static std::atomic<bool> g_tune_init{false};
static const char* const kTuneNames[TUNE_COUNT] = {"epi_l2_prefetch", "tc_resident", "tcp_epi_warps", "tc_form_c64", "tc_pair",
                                                   "wait_backoff_ns", "pdl", "wgrad_multicast", "tct_band", "tct_debug", "tct_products", "mma_warp_high", "wgrad64_products", "mnist_fused", "uniform_issue", "gn_block", "tcp2_half_stage", "tcp2_halo", "wgrad_htaps", "peer_form"};
static const char* const kTuneEnv[TUNE_COUNT] = {"MSB_EPI_L2_PREFETCH", "MSB_TC_RESIDENT", "MSB_TCP_EPI_WARPS", "MSB_TC_FORM_C64",
                                                 "MSB_TC_PAIR", "MSB_WAIT_BACKOFF_NS", "MSB_PDL", "MSB_WGRAD_MULTICAST", "MSB_TCT_BAND", "MSB_TCT_DEBUG", "MSB_TCT_PRODUCTS", "MSB_MMA_WARP_HIGH", "MSB_WGRAD64_PRODUCTS", "MSB_MNIST_FUSED", "MSB_UNIFORM_ISSUE", "MSB_GN_BLOCK", "MSB_TCP2_HALF_STAGE", "MSB_TCP2_HALO", "MSB_WGRAD_HTAPS", "MSB_PEER_FORM"};
static const int kTuneDefault[TUNE_COUNT] = {0, 0, 16, 2, 2, 0, 1, 0, 0, 0, 4, 0, 4, 1, 1, 1, 0, 1, 1, 0};
static void tune_init() {
    if (g_tune_init.load(std::memory_order_acquire)) return;
    for (int i = 0; i < TUNE_COUNT; ++i) {
        const char* e = getenv(kTuneEnv[i]);
        g_tune[i].store(e ? atoi(e) : kTuneDefault[i]);
    }
    g_tune_init.store(true, std::memory_order_release);
}
int tune_get(int which) { tune_init(); return g_tune[which].load(std::memory_order_relaxed); }
int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

// ---- optional per-kernel timing (bench.py roofline): CUDA events around engine launches ----
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev;      // pairs (start, stop)
    std::vector<int> kind;            // kind of each pair
    std::vector<double> flops;        // algorithmic flops of each pair
    std::vector<double> exec;         // executed tensor-core flops of each pair (hi/lo products formed)
    size_t used = 0;
};
static Prof g_prof;
static int prof_begin(int kind, double flops, double products, cudaStream_t st) {
    if (!g_prof.on) return -1;
    if (g_prof.used + 2 > g_prof.ev.size()) {
        for (int i = 0; i < 512; ++i) { cudaEvent_t e; cudaEventCreate(&e); g_prof.ev.push_back(e); }
    }
    int id = (int)(g_prof.used / 2);
    g_prof.kind.resize(id + 1); g_prof.flops.resize(id + 1); g_prof.exec.resize(id + 1);
    g_prof.kind[id] = kind; g_prof.flops[id] = flops; g_prof.exec[id] = flops * products;
    cudaEventRecord(g_prof.ev[g_prof.used], st);
    g_prof.used += 2;
    return id;
}
static void prof_end(int id, cudaStream_t st) {
    if (id >= 0) cudaEventRecord(g_prof.ev[2 * id + 1], st);
}

// ---- host helpers shared with blocks.cu (declared in msb_host.h) ----
// Which form of the tcgen05 convolution runs for a shape.  Measured on B200 (profiles/): the pixel-major form (3 hi/lo
// products, CTA pair) wins for C = 128; for C = 64 its N = 128 / N = 64 MMAs are shared-memory-operand-bound, and the form
// with the weights resident in tensor memory (conv_tct.cu: A from TMEM, only the activations read from shared memory)
// is the default wherever it tiles (32-pixel-wide images), the pixel-major form elsewhere.
// MSB_TC_CONV=cm|pm forces one of the older forms for every shape (A/B runs).
int tc_form(int C, int H, int W) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("MSB_TC_CONV");
        forced = !e ? -1 : (strcmp(e, "cm") == 0 ? 0 : (strcmp(e, "pm") == 0 ? 1 : -1));
    }
    if (forced >= 0) return forced;
    if (C >= 128) return 1;
    const int v = tune_get(TUNE_TC_FORM_C64);
    if (v == 2) return tct_shape_supported(C, H, W) ? 2 : 1;
    return v == 0 ? 0 : 1;
}
size_t packed_w_bytes(int engine, int C) {
    if (engine != MSB_ENGINE_TCGEN05) return (size_t)9 * C * C * sizeof(float);
    size_t n = std::max(tcp_packed_weight_bytes(C), tc_packed_weight_bytes(C));     // any form
    if (C == 64) n = std::max(n, tct_packed_weight_bytes());
    return n;
}
int resolve_engine_shape(int engine, int C, int H, int W) {
    const bool tc_ok = tc_shape_supported(C, H, W);
    if (engine == MSB_ENGINE_SIMT) return MSB_ENGINE_SIMT;
    if (engine == MSB_ENGINE_TCGEN05) {
        if (!tc_ok) { set_error("tcgen05 engine does not cover C=%d H=%d W=%d", C, H, W); return -1; }
        return MSB_ENGINE_TCGEN05;
    }
    if (engine == MSB_ENGINE_AUTO) return tc_ok ? MSB_ENGINE_TCGEN05 : MSB_ENGINE_SIMT;
    set_error("unknown engine %d", engine);
    return -1;
}
double conv_flops(ConvShape s) { return 2.0 * s.B * s.H * s.W * (double)s.C * 9.0 * s.C; }
int run_conv(int engine, const __nv_bfloat16* in, const void* wpacked, const EpiParams& e, ConvShape s, cudaStream_t st) {
    const int form = engine == MSB_ENGINE_TCGEN05 ? tc_form(s.C, s.H, s.W) : -1;
    const double products = form < 0 ? 1.0 : (form == 1 ? 3.0 : (form == 2 ? (double)tct_products() : 4.0));
    int id = prof_begin(MSB_PROF_CONV, conv_flops(s), products, st);
    int rc;
    if (form == 2) rc = launch_conv3x3_tct(in, wpacked, e, s, st);
    else if (form == 1)
        rc = ((tune_get(TUNE_TC_PAIR) == 1 || (tune_get(TUNE_TC_PAIR) == 2 && s.C >= 128)) && tcp2_shape_supported(s.B, s.C, s.H, s.W))
                 ? launch_conv3x3_tcp2(in, (const __nv_bfloat16*)wpacked, e, s, st)
                 : launch_conv3x3_tcp(in, (const __nv_bfloat16*)wpacked, e, s, st);
    else if (form == 0) rc = launch_conv3x3_tc(in, (const __nv_bfloat16*)wpacked, e, s, st);
    else rc = launch_conv3x3_simt(in, (const float*)wpacked, e, s, st);
    prof_end(id, st);
    return rc;
}
void pack_w(int engine, const float* w, void* out, int C, int transpose, cudaStream_t st, int H, int W) {
    if (engine == MSB_ENGINE_TCGEN05) {
        const int form = tc_form(C, H, W);
        if (form == 2) launch_pack_w_tct(w, out, transpose, st);
        else if (form == 1) launch_pack_w_tcp(w, (__nv_bfloat16*)out, C, transpose, st);
        else launch_pack_w_tc(w, (__nv_bfloat16*)out, C, transpose, st);
    } else launch_pack_w_simt(w, (float*)out, C, C, 0, transpose, st);
}
bool use_tc_wgrad(int engine, ConvShape s) { return engine == MSB_ENGINE_TCGEN05 && wgrad_tc_supported(s); }
int wgrad_nparts(int engine, ConvShape s) {
    return use_tc_wgrad(engine, s) ? wgrad_tc_nparts(s) : wgrad_simt_nparts(s);
}
// Weight-gradient accumulation over the launches of one backward pass.
//  tcgen05: every launch adds onto the CTA-private partial slots (first launch overwrites); ONE
//           fixed-order reduction per weight tensor at the end (wgrad_finish).
//  SIMT   : partials are reduced into grad_w after every launch.
int run_wgrad(int engine, const __nv_bfloat16* gout, const __nv_bfloat16* in, WgradAcc& acc, ConvShape s, cudaStream_t st) {
    int nparts = 0, rc;
    const bool tc = use_tc_wgrad(engine, s);
    int id = prof_begin(MSB_PROF_WGRAD, conv_flops(s), tc ? (s.C == 64 ? (tune_get(TUNE_WGRAD64_PRODUCTS) == 3 ? 10.0 / 3.0 : 4.0) : 3.0) : 1.0, st);   // C = 64: two taps at 3 products, the unpaired third at 4
    if (tc) rc = launch_wgrad3x3_tc(gout, in, acc.partial, &nparts, acc.launches > 0, s, st);
    else rc = launch_wgrad3x3_simt(gout, in, acc.partial, &nparts, s, st);
    prof_end(id, st);
    if (rc) return rc;
    acc.nparts = nparts;
    if (!tc) launch_wgrad_reduce(acc.partial, nparts, acc.grad_w, s.C, acc.launches > 0, st);
    acc.launches++;
    return check_cuda(cudaGetLastError(), "wgrad launch");
}
int wgrad_finish(int engine, WgradAcc& acc, ConvShape s, cudaStream_t st) {
    if (use_tc_wgrad(engine, s) && acc.launches > 0) launch_wgrad_reduce(acc.partial, acc.nparts, acc.grad_w, s.C, 0, st);
    return check_cuda(cudaGetLastError(), "wgrad reduce launch");
}


namespace {

int validate(const MsbOdeDesc* d) {
    if (!d) { set_error("null descriptor"); return -1; }
    if (d->rhs_kind != MSB_RHS_PREACT_NF && d->rhs_kind != MSB_RHS_POSTACT_NF && d->rhs_kind != MSB_RHS_MNIST_GN_T &&
        d->rhs_kind != MSB_RHS_PREACT_GN && d->rhs_kind != MSB_RHS_POSTACT_GN) {
        set_error("rhs_kind %d is not implemented (supported: PREACT_NF, POSTACT_NF, MNIST_GN_T, PREACT_GN, POSTACT_GN)", d->rhs_kind);
        return -1;
    }
    if (d->act != MSB_ACT_GELU_ERF && d->act != MSB_ACT_RELU && d->act != MSB_ACT_NONE) {
        set_error("unsupported activation %d", d->act); return -1;
    }
    if (d->stages < 1 || d->stages > MSB_MAX_STAGES) { set_error("stages must be 1..4 (got %d)", d->stages); return -1; }
    if (d->n_steps < 1) { set_error("n_steps must be >= 1 (got %d)", d->n_steps); return -1; }
    if (d->batch < 1 || d->height < 1 || d->width < 1 || d->channels < 4 || d->channels % 4) {
        set_error("bad shape B=%d H=%d W=%d C=%d (C must be a positive multiple of 4)", d->batch, d->height, d->width, d->channels);
        return -1;
    }
    if (!d->time_grid) { set_error("time_grid is NULL"); return -1; }
    if (d->n_solvers < 0 || d->n_solvers > kMaxSlices) { set_error("n_solvers must be 0..%d (got %d)", kMaxSlices, d->n_solvers); return -1; }
    if (d->n_solvers > 1) {
        if (!d->solver_tableaus) { set_error("n_solvers = %d but solver_tableaus is NULL", d->n_solvers); return -1; }
        if (d->batch % d->n_solvers) { set_error("batch %d is not divisible into %d solver slices", d->batch, d->n_solvers); return -1; }
        if (d->rhs_kind == MSB_RHS_MNIST_GN_T || d->rhs_kind == MSB_RHS_PREACT_GN || d->rhs_kind == MSB_RHS_POSTACT_GN) {
            set_error("the stacked solver axis is not implemented for the GroupNorm right-hand sides");
            return -1;
        }
    }
    return 0;
}

// Tableau of every solver slice (one entry when there is no stacked solver axis).
struct Tabs {
    int K;
    MsbTableau t[kMaxSlices];
    int slice_batch;     // images per slice, 0 when K == 1 (EpiParams::slice_batch)
};
Tabs make_tabs(const MsbOdeDesc* d) {
    Tabs r;
    r.K = d->n_solvers > 1 ? d->n_solvers : 1;
    if (r.K == 1) {
        memcpy(r.t[0].c, d->c, sizeof(d->c)); memcpy(r.t[0].b, d->b, sizeof(d->b)); memcpy(r.t[0].w, d->w, sizeof(d->w));
        r.slice_batch = 0;
    } else {
        for (int s = 0; s < r.K; ++s) r.t[s] = d->solver_tableaus[s];
        r.slice_batch = d->batch / r.K;
    }
    return r;
}

// Resolve MSB_ENGINE_AUTO.  Both engines are this library's own CUDA kernels; there is no
// library (cuDNN) or CPU path to fall back to.
int resolve_engine(const MsbOdeDesc* d) {
    bool tc_ok = tc_shape_supported(d->channels, d->height, d->width);
    if (d->rhs_kind == MSB_RHS_MNIST_GN_T) {
        if (d->engine == MSB_ENGINE_TCGEN05) { set_error("the MNIST right-hand side runs on the SIMT engine only"); return -1; }
        return MSB_ENGINE_SIMT;
    }
    (void)tc_ok;
    return resolve_engine_shape(d->engine, d->channels, d->height, d->width);
}

size_t state_elems(const MsbOdeDesc* d) { return (size_t)d->batch * d->height * d->width * d->channels; }

struct TapeSlot { __nv_bfloat16* A; float* G0; __nv_bfloat16* Hs; float* G1; };
// the same slot, advanced to the images [b0, ...) of the batch (`off` = b0 * H*W*C state elements)
TapeSlot slot_at(TapeSlot t, size_t off) {
    t.A += 2 * off; t.Hs += 2 * off;              // split tensors hold two bf16 planes per state element
    if (t.G0) t.G0 += off;
    if (t.G1) t.G1 += off;
    return t;
}
// Micro-batching (L2 locality): an ODE block is a chain of launches in which every tensor is consumed by the launch
// right after the one that wrote it.  At B = 512 a state tensor is 134 MB (> the 126 MB L2), so every hand-off
// round-trips HBM.  Integrating the batch in slices of `MSB_MICROBATCH` images, each slice taken through all steps
// before the next one starts, keeps those hand-offs L2-resident; samples are independent, results are unchanged.
int microbatch_images(const MsbOdeDesc* d, int n_slices) {
    static int env = -2;
    if (env == -2) { const char* e = getenv("MSB_MICROBATCH"); env = e ? atoi(e) : -1; }
    if (n_slices > 1) return d->batch;              // stacked solver axis: one launch covers every slice
    int mb = env;
    if (mb < 0) mb = 0;                              // default: off (see DESIGN.md for the measured trade-off)
    if (mb <= 0 || mb >= d->batch || d->batch % mb) return d->batch;   // equal slices only (wgrad partial slots)
    return mb;
}
TapeSlot tape_slot(void* tape, size_t E, int slot) {
    char* p = (char*)tape + (size_t)slot * 4 * align_up(E * 4);
    size_t q = align_up(E * 4);
    return TapeSlot{(__nv_bfloat16*)p, (float*)(p + q), (__nv_bfloat16*)(p + 2 * q), (float*)(p + 3 * q)};
}

}  // namespace
}  // namespace msb

using namespace msb;

extern "C" {

int msb_abi_version(void) { return MSB_ABI_VERSION; }
size_t msb_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(MsbOdeDesc);
        case 1: return sizeof(MsbTableau);
        case 2: return sizeof(MsbMnistParams);
        case 3: return sizeof(MsbMnistGrads);
        case 4: return sizeof(MsbDownDesc);
        default: return 0;
    }
}
const char* msb_last_error(void) { return g_err.c_str(); }
uint64_t msb_launch_count(void) { return g_launches.load(); }
int msb_set_option(const char* name, int value) {
    if (!name) { set_error("msb_set_option: null name"); return -1; }
    tune_init();
    for (int i = 0; i < TUNE_COUNT; ++i)
        if (strcmp(name, kTuneNames[i]) == 0) { g_tune[i].store(value); return 0; }
    set_error("msb_set_option: unknown option '%s'", name);
    return -1;
}
int msb_get_option(const char* name, int* value) {
    if (!name || !value) { set_error("msb_get_option: null argument"); return -1; }
    for (int i = 0; i < TUNE_COUNT; ++i)
        if (strcmp(name, kTuneNames[i]) == 0) { *value = tune_get(i); return 0; }
    set_error("msb_get_option: unknown option '%s'", name);
    return -1;
}

int msb_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.used = 0;
    return 0;
}
int msb_profile_read(int kind, double* total_ms, double* total_flops, int64_t* count) {
    double ms = 0, fl = 0; int64_t n = 0;
    for (size_t id = 0; id < g_prof.used / 2; ++id) {
        if (g_prof.kind[id] != kind) continue;
        if (cudaEventSynchronize(g_prof.ev[2 * id + 1]) != cudaSuccess) { set_error("profile event sync failed"); return -1; }
        float t = 0.f;
        if (cudaEventElapsedTime(&t, g_prof.ev[2 * id], g_prof.ev[2 * id + 1]) != cudaSuccess) { set_error("profile elapsed failed"); return -1; }
        ms += t; fl += g_prof.flops[id]; ++n;
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (count) *count = n;
    return 0;
}

int msb_profile_read_executed(int kind, double* executed_flops) {
    double fl = 0;
    for (size_t id = 0; id < g_prof.used / 2; ++id)
        if (g_prof.kind[id] == kind) fl += g_prof.exec[id];
    if (executed_flops) *executed_flops = fl;
    return 0;
}

int msb_device_supports_tcgen05(int device) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
        set_error("cannot query device %d", device);
        return -1;
    }
    return major == 10 ? 1 : 0;
}
int msb_shape_supports_tcgen05(int channels, int height, int width) {
    return tc_shape_supported(channels, height, width) ? 1 : 0;
}

size_t msb_odeblock_workspace_bytes(const MsbOdeDesc* d) {
    if (validate(d)) return 0;
    int engine = resolve_engine(d);
    if (engine < 0) return 0;
    size_t E = state_elems(d);
    size_t n = 0;
    n += 2 * align_up(packed_w_bytes(engine, d->channels));
    n += 2 * align_up(E * 4);                                  // y ping-pong
    n += (size_t)(d->stages - 1) * align_up(E * 4);            // k_1 .. k_{s-1}
    n += 2 * align_up(E * 4);                                  // A / Hs split (inference)
    if (d->rhs_kind == MSB_RHS_MNIST_GN_T)                     // conv output, stage input, 2 tapmaps; fused path: packed weights + maps
        n += 2 * align_up(E * 4) + 2 * align_up((size_t)d->height * d->width * d->channels * 4) + mnist_fused_workspace_bytes() + 1024;
    if (d->rhs_kind == MSB_RHS_PREACT_GN) n += 2 * align_up(E * 4);     // conv1 output, stage input
    if (d->rhs_kind == MSB_RHS_POSTACT_GN) n += 3 * align_up(E * 4);    // conv1 / conv2 outputs, stage input
    return n + 4096;
}
size_t msb_odeblock_tape_bytes(const MsbOdeDesc* d) {
    if (validate(d)) return 0;
    const int per_slot = d->rhs_kind == MSB_RHS_MNIST_GN_T ? 5 : 4;      // MNIST: X, P1, P2, A, Hs;  PREACT_GN: X, P1, A, Hs;  POSTACT_GN: P1, P2, A, Hs
    return (size_t)d->n_steps * d->stages * per_slot * align_up(state_elems(d) * 4);
}
size_t msb_odeblock_bwd_workspace_bytes(const MsbOdeDesc* d) {
    if (validate(d)) return 0;
    int engine = resolve_engine(d);
    if (engine < 0) return 0;
    size_t E = state_elems(d);
    if (d->rhs_kind == MSB_RHS_MNIST_GN_T) {
        ConvShape s{d->batch, d->height, d->width, d->channels};
        size_t n = 2 * align_up((size_t)9 * d->channels * d->channels * 4);      // transposed packed weights
        n += 2 * align_up(E * 4);                                                // gbar ping-pong
        n += (size_t)(d->stages - 1) * align_up(E * 4);                          // xbar_1 .. xbar_{s-1}
        n += 4 * align_up(E * 4);                                                // kbar, dP, dH (fp32), Dsplit
        n += align_up((size_t)wgrad_simt_nparts(s) * 9 * d->channels * d->channels * 4);
        n += 6 * align_up((size_t)d->batch * d->channels * 4);                   // per-sample dgamma / dbeta partials
        return n + 4096;
    }
    ConvShape s{d->batch, d->height, d->width, d->channels};
    size_t n = 0;
    n += 2 * align_up(packed_w_bytes(engine, d->channels));
    n += 2 * align_up(E * 4);                                  // gbar ping-pong
    n += (size_t)(d->stages - 1) * align_up(E * 4);            // xbar_1 .. xbar_{s-1}
    n += 2 * align_up(E * 4);                                  // Kbar / DP split
    n += 2 * align_up((size_t)wgrad_nparts(engine, s) * 9 * d->channels * d->channels * 4);
    if (d->rhs_kind == MSB_RHS_PREACT_GN || d->rhs_kind == MSB_RHS_POSTACT_GN)     // dH (fp32) + per-sample dgamma / dbeta partials of 2 norms
        n += align_up(E * 4) + 4 * align_up((size_t)d->batch * d->channels * 4);
    return n + 4096;
}

size_t msb_odeblock_bwd_workspace_bytes_tableau(const MsbOdeDesc* d) {
    const size_t base = msb_odeblock_bwd_workspace_bytes(d);
    if (base == 0) return 0;
    if (d->rhs_kind == MSB_RHS_MNIST_GN_T)     // recomputed k_1..k_S, the two time-channel tap maps, the reduction scratch
        return base + (size_t)d->stages * align_up(state_elems(d) * 4) +
               2 * align_up((size_t)d->height * d->width * d->channels * 4) + align_up(dot_scratch_bytes()) + 1024;
    int engine = resolve_engine(d);
    // forward-packed conv2 weights, the recomputed stage derivatives k_1..k_S, the reduction scratch
    return base + align_up(packed_w_bytes(engine, d->channels)) + (size_t)d->stages * align_up(state_elems(d) * 4) +
           align_up(dot_scratch_bytes()) + 1024;
}


// ---------------------------------------------------------------------------------------------
// MNIST right-hand side (forward):  f(t, x) = GN3(cconv2(t, relu(GN2(cconv1(t, relu(GN1(x)))))))
// sopa/src/models/odenet_mnist/layers.py:158-171.  Five launches per evaluation; the RK stage
// combination is the epilogue of the GN3 launch.  Stage time t_i = t_n + c_i*dt (`_get_t`).
// ---------------------------------------------------------------------------------------------
struct MnistSlot { float *X, *P1, *P2; __nv_bfloat16 *A, *Hs; };
static MnistSlot mnist_slot(void* tape, size_t E, int slot) {
    const size_t q = align_up(E * 4);
    char* p = (char*)tape + (size_t)slot * 5 * q;
    return MnistSlot{(float*)p, (float*)(p + q), (float*)(p + 2 * q), (__nv_bfloat16*)(p + 3 * q), (__nv_bfloat16*)(p + 4 * q)};
}
// stage time t_n + c_i*dt with the reference's two roundings (`_get_t`)
static float mnist_stage_time(const MsbOdeDesc* d, int n, int i) {
    const float t0 = d->time_grid[n];
    const float dt = d->time_grid[n + 1] - t0;
    volatile float cdt = d->c[i] * dt;
    return (i == 0) ? t0 : t0 + cdt;
}

static int mnist_forward(const MsbOdeDesc* d, const float* x, const MsbMnistParams* mp, float* y_out, void* workspace,
                         size_t workspace_bytes, void* tape, size_t tape_bytes, cudaStream_t st) {
    const bool save = d->save_tape != 0;
    if (save && (!tape || tape_bytes < msb_odeblock_tape_bytes(d))) { set_error("tape missing or too small"); return -1; }
    if (!mp) { set_error("MSB_RHS_MNIST_GN_T needs MsbMnistParams"); return -1; }
    for (int i = 0; i < 3; ++i) if (!mp->norm_w[i] || !mp->norm_b[i]) { set_error("MNIST params: null norm pointer"); return -1; }
    for (int i = 0; i < 2; ++i) if (!mp->conv_w[i] || !mp->conv_b[i]) { set_error("MNIST params: null conv pointer"); return -1; }
    if (!x || !y_out || !workspace) { set_error("null pointer argument"); return -1; }
    if (workspace_bytes < msb_odeblock_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    if (mnist_fused_supported(d, mp))       // the whole solve in ONE persistent tcgen05 launch (mnist_fused.cu)
        return launch_mnist_fused_forward(d, x, mp, y_out, workspace, save ? tape : nullptr, align_up(E * 4), st);
    ConvShape shp{d->batch, d->height, d->width, C};
    Carver cv(workspace, workspace_bytes);
    float* wp[2] = {cv.take<float>((size_t)9 * C * C * 4), cv.take<float>((size_t)9 * C * C * 4)};
    float* ybuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* kbuf[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < S - 1; ++i) kbuf[i] = cv.take<float>(E * 4);
    __nv_bfloat16* A = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* Hs = cv.take<__nv_bfloat16>(E * 4);
    float* PQ = cv.take<float>(E * 4);
    float* xbuf = cv.take<float>(E * 4);
    float* tapmap[2] = {cv.take<float>((size_t)d->height * d->width * C * 4), cv.take<float>((size_t)d->height * d->width * C * 4)};
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    for (int k = 0; k < 2; ++k) {
        launch_pack_w_simt(mp->conv_w[k], wp[k], C, C + 1, 1, 0, st);
        launch_time_tapmap(mp->conv_w[k], tapmap[k], d->height, d->width, C, st);
    }
    // With a tape every intermediate a backward pass needs is written straight into its slot: the stage input X,
    // the two convolution outputs P1 / P2 (GroupNorm inputs) and the two convolution operands A / Hs.
    const float* y_cur = x;
    if (save) {
        MnistSlot s0 = mnist_slot(tape, E, 0);
        if (check_cuda(cudaMemcpyAsync(s0.X, x, E * 4, cudaMemcpyDeviceToDevice, st), "copy x to tape")) return -1;
        y_cur = s0.X;
    }
    for (int n = 0; n < N; ++n) {
        const float dt = d->time_grid[n + 1] - d->time_grid[n];
        float* y_next = (n == N - 1) ? y_out : (save ? mnist_slot(tape, E, (n + 1) * S).X : ybuf[n & 1]);
        for (int i = 0; i < S; ++i) {
            const float ti = mnist_stage_time(d, n, i);
            MnistSlot sl = save ? mnist_slot(tape, E, n * S + i) : MnistSlot{nullptr, PQ, PQ, A, Hs};
            const float* xi = (i == 0) ? y_cur : (save ? sl.X : xbuf);
            EpiParams g1 = epi_default();
            g1.act = ACT_RELU; g1.out_split = sl.A;
            if (launch_groupnorm_epi(xi, mp->norm_w[0], mp->norm_b[0], g1, shp, mp->groups, mp->eps, st)) return -1;
            EpiParams c1 = epi_default();
            c1.chan_bias = mp->conv_b[0]; c1.pix_bias = tapmap[0]; c1.pix_bias_scale = ti; c1.out_f32 = sl.P1;
            if (run_conv(MSB_ENGINE_SIMT, sl.A, wp[0], c1, shp, st)) return -1;
            EpiParams g2 = epi_default();
            g2.act = ACT_RELU; g2.out_split = sl.Hs;
            if (launch_groupnorm_epi(sl.P1, mp->norm_w[1], mp->norm_b[1], g2, shp, mp->groups, mp->eps, st)) return -1;
            EpiParams c2 = epi_default();
            c2.chan_bias = mp->conv_b[1]; c2.pix_bias = tapmap[1]; c2.pix_bias_scale = ti; c2.out_f32 = sl.P2;
            if (run_conv(MSB_ENGINE_SIMT, sl.Hs, wp[1], c2, shp, st)) return -1;
            float* const PQ2 = sl.P2;
            EpiParams g3 = epi_default();                       // k_i = GN3(.) and the RK combination
            g3.base = y_cur; g3.k[0].dt = dt;
            if (i < S - 1) {
                g3.v_out = kbuf[i];
                g3.nsrc = i;
                for (int j = 0; j < i; ++j) { g3.src[j] = kbuf[j]; g3.k[0].coef[j] = d->w[(i + 1) * MSB_MAX_STAGES + j]; }
                g3.k[0].coef_v = d->w[(i + 1) * MSB_MAX_STAGES + i];
                g3.out_f32 = save ? mnist_slot(tape, E, n * S + i + 1).X : xbuf;
            } else {
                g3.nsrc = S - 1;
                for (int j = 0; j < S - 1; ++j) { g3.src[j] = kbuf[j]; g3.k[0].coef[j] = d->b[j]; }
                g3.k[0].coef_v = d->b[S - 1];
                g3.out_f32 = y_next;
            }
            if (launch_groupnorm_epi(PQ2, mp->norm_w[2], mp->norm_b[2], g3, shp, mp->groups, mp->eps, st)) return -1;
        }
        y_cur = y_next;
    }
    return check_cuda(cudaGetLastError(), "odeblock forward (mnist)");
}


// ---------------------------------------------------------------------------------------------
// CIFAR pre-activation right-hand side WITH GroupNorm (the 'GN' / 'LN' / 'IN' normalisations of
// sopa/src/models/odenet_cifar10/utils.py:26-36 inside PreBasicBlock2, layers.py:148-161):
//     f(x) = conv2(act(GN2(conv1(act(GN1(x))))))
// A normalisation needs whole-image statistics of a convolution OUTPUT, so it cannot ride in that convolution's tile
// epilogue: each GroupNorm (+ activation + hi/lo split) is its own launch between the convolutions; the Runge-Kutta
// stage combination stays fused in conv2's epilogue.  Tape per (step, stage): X (stage input), P1 (conv1 output),
// A = split(act(GN1(X))), Hs = split(act(GN2(P1))).  Convolutions run on the engine of the shape (tcgen05 where tiled).
// ---------------------------------------------------------------------------------------------
struct GnSlot { float *X, *P1; __nv_bfloat16 *A, *Hs; };
static GnSlot gn_slot(void* tape, size_t E, int slot) {
    const size_t q = align_up(E * 4);
    char* p = (char*)tape + (size_t)slot * 4 * q;
    return GnSlot{(float*)p, (float*)(p + q), (__nv_bfloat16*)(p + 2 * q), (__nv_bfloat16*)(p + 3 * q)};
}
static int gn_check_params(const MsbMnistParams* mp) {
    if (!mp) { set_error("MSB_RHS_PREACT_GN needs its parameters in MsbMnistParams"); return -1; }
    for (int i = 0; i < 2; ++i)
        if (!mp->norm_w[i] || !mp->norm_b[i] || !mp->conv_w[i]) { set_error("PREACT_GN params: null norm / conv pointer"); return -1; }
    return 0;
}

static int gn_preact_forward(const MsbOdeDesc* d, const float* x, const MsbMnistParams* mp, float* y_out, void* workspace,
                             size_t workspace_bytes, void* tape, size_t tape_bytes, cudaStream_t st) {
    const bool save = d->save_tape != 0;
    if (save && (!tape || tape_bytes < msb_odeblock_tape_bytes(d))) { set_error("tape missing or too small"); return -1; }
    if (gn_check_params(mp)) return -1;
    if (!x || !y_out || !workspace) { set_error("null pointer argument"); return -1; }
    if (workspace_bytes < msb_odeblock_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const int engine = resolve_engine(d);
    if (engine < 0) return -1;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    ConvShape shp{d->batch, d->height, d->width, C};
    Carver cv(workspace, workspace_bytes);
    void* wp[2] = {cv.take<char>(packed_w_bytes(engine, C)), cv.take<char>(packed_w_bytes(engine, C))};
    float* ybuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* kbuf[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < S - 1; ++i) kbuf[i] = cv.take<float>(E * 4);
    __nv_bfloat16* A = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* Hs = cv.take<__nv_bfloat16>(E * 4);
    float* P = cv.take<float>(E * 4);
    float* xbuf = cv.take<float>(E * 4);
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    for (int k = 0; k < 2; ++k) pack_w(engine, mp->conv_w[k], wp[k], C, 0, st, d->height, d->width);
    const float* y_cur = x;
    if (save) {
        GnSlot s0 = gn_slot(tape, E, 0);
        if (check_cuda(cudaMemcpyAsync(s0.X, x, E * 4, cudaMemcpyDeviceToDevice, st), "copy x to tape")) return -1;
        y_cur = s0.X;
    }
    for (int n = 0; n < N; ++n) {
        const float dt = d->time_grid[n + 1] - d->time_grid[n];
        float* y_next = (n == N - 1) ? y_out : (save ? gn_slot(tape, E, (n + 1) * S).X : ybuf[n & 1]);
        for (int i = 0; i < S; ++i) {
            GnSlot sl = save ? gn_slot(tape, E, n * S + i) : GnSlot{nullptr, P, A, Hs};
            const float* xi = (i == 0) ? y_cur : (save ? sl.X : xbuf);
            EpiParams g1 = epi_default();
            g1.act = d->act; g1.out_split = sl.A;
            if (launch_groupnorm_epi(xi, mp->norm_w[0], mp->norm_b[0], g1, shp, mp->groups, mp->eps, st)) return -1;
            EpiParams c1 = epi_default();
            c1.out_f32 = sl.P1;
            if (run_conv(engine, sl.A, wp[0], c1, shp, st)) return -1;
            EpiParams g2 = epi_default();
            g2.act = d->act; g2.out_split = sl.Hs;
            if (launch_groupnorm_epi(sl.P1, mp->norm_w[1], mp->norm_b[1], g2, shp, mp->groups, mp->eps, st)) return -1;
            EpiParams e2 = epi_default();                       // k_i = conv2(.) and the RK combination
            e2.base = y_cur; e2.k[0].dt = dt;
            if (i < S - 1) {
                e2.v_out = kbuf[i];
                e2.nsrc = i;
                for (int j = 0; j < i; ++j) { e2.src[j] = kbuf[j]; e2.k[0].coef[j] = d->w[(i + 1) * MSB_MAX_STAGES + j]; }
                e2.k[0].coef_v = d->w[(i + 1) * MSB_MAX_STAGES + i];
                e2.out_f32 = save ? gn_slot(tape, E, n * S + i + 1).X : xbuf;
            } else {
                e2.nsrc = S - 1;
                for (int j = 0; j < S - 1; ++j) { e2.src[j] = kbuf[j]; e2.k[0].coef[j] = d->b[j]; }
                e2.k[0].coef_v = d->b[S - 1];
                e2.out_f32 = y_next;
            }
            if (run_conv(engine, sl.Hs, wp[1], e2, shp, st)) return -1;
        }
        y_cur = y_next;
    }
    return check_cuda(cudaGetLastError(), "odeblock forward (preact GN)");
}

//   kbar_i --conv2^T--> dH --act', GN2'--> dP1 --conv1^T--> dH --act', GN1'--> xbar_i  (+ the RK adjoint combination)
static int gn_preact_backward(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mp, const void* tape,
                              size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads, void* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
    if (gn_check_params(mp)) return -1;
    if (!grad_y || !tape || !grad_x || !workspace) { set_error("null pointer argument"); return -1; }
    if (tape_bytes < msb_odeblock_tape_bytes(d)) { set_error("tape too small"); return -1; }
    if (workspace_bytes < msb_odeblock_bwd_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const bool need_w = grads != nullptr;
    if (need_w)
        for (int i = 0; i < 2; ++i)
            if (!grads->norm_w[i] || !grads->norm_b[i] || !grads->conv_w[i]) { set_error("PREACT_GN grads: null pointer"); return -1; }
    const int engine = resolve_engine(d);
    if (engine < 0) return -1;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    ConvShape shp{d->batch, d->height, d->width, C};
    Carver cv(workspace, workspace_bytes);
    void* wt[2] = {cv.take<char>(packed_w_bytes(engine, C)), cv.take<char>(packed_w_bytes(engine, C))};
    float* gbuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* xbar[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 1; i < S; ++i) xbar[i] = cv.take<float>(E * 4);
    __nv_bfloat16* Kbar = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* DP = cv.take<__nv_bfloat16>(E * 4);
    const size_t part_bytes = (size_t)wgrad_nparts(engine, shp) * 9 * C * C * 4;
    WgradAcc acc1{cv.take<float>(part_bytes), need_w ? grads->conv_w[0] : nullptr, 0, 0};
    WgradAcc acc2{cv.take<float>(part_bytes), need_w ? grads->conv_w[1] : nullptr, 0, 0};
    float* dH = cv.take<float>(E * 4);
    float* gnpart[4];
    for (int i = 0; i < 4; ++i) gnpart[i] = cv.take<float>((size_t)d->batch * C * 4);     // (dgamma_k, dbeta_k), k = 1, 2
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    for (int k = 0; k < 2; ++k) pack_w(engine, mp->conv_w[k], wt[k], C, 1, st, d->height, d->width);

    auto dt_of = [&](int n) { return d->time_grid[n + 1] - d->time_grid[n]; };
    launch_act_split(grad_y, nullptr, ACT_NONE, dt_of(N - 1) * d->b[S - 1], Kbar, nullptr, d->batch, d->height, d->width, C, st);
    int evals = 0;
    const float* g_cur = grad_y;
    for (int n = N - 1; n >= 0; --n) {
        const float dt = dt_of(n);
        float* g_next = (n == 0) ? grad_x : gbuf[n & 1];
        for (int i = S - 1; i >= 0; --i) {
            const GnSlot sl = gn_slot(const_cast<void*>(tape), E, n * S + i);
            const int acc = evals > 0;
            // Kbar = split(kbar_i).   dW2 += kbar_i (x) Hs_i ;  dH = dgrad_W2(kbar_i)
            if (need_w && run_wgrad(engine, Kbar, sl.Hs, acc2, shp, st)) return -1;
            EpiParams e = epi_default();
            e.out_f32 = dH;
            if (run_conv(engine, Kbar, wt[1], e, shp, st)) return -1;
            // dP1 = GN2'(act'(.) dH)
            e = epi_default();
            e.out_split = DP;
            if (launch_groupnorm_bwd_epi(sl.P1, mp->norm_w[1], mp->norm_b[1], dH, 1.f, d->act, e, need_w ? gnpart[2] : nullptr,
                                         need_w ? gnpart[3] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            // dW1 += dP1 (x) A_i ;  dH = dgrad_W1(dP1)
            if (need_w && run_wgrad(engine, DP, sl.A, acc1, shp, st)) return -1;
            e = epi_default();
            e.out_f32 = dH;
            if (run_conv(engine, DP, wt[0], e, shp, st)) return -1;
            // xbar_i = GN1'(act'(.) dH), then the adjoint stage combination (as in the normalisation-free path)
            EpiParams e4 = epi_default();
            e4.base = g_cur;
            if (i > 0) {
                e4.v_out = xbar[i];
                e4.base_is_one = 0;
                e4.k[0].base_coef = dt * d->b[i - 1];
                int ns = 0;
                for (int j = S - 1; j > i; --j) { e4.src[ns] = xbar[j]; e4.k[0].coef[ns] = d->w[j * MSB_MAX_STAGES + (i - 1)]; ++ns; }
                e4.nsrc = ns;
                e4.k[0].coef_v = d->w[i * MSB_MAX_STAGES + (i - 1)];
                e4.k[0].dt = dt;
                e4.out_split = Kbar;                                  // = split(kbar_{i-1})
            } else {
                int ns = 0;
                for (int j = S - 1; j > 0; --j) { e4.src[ns] = xbar[j]; e4.k[0].coef[ns] = 1.f; ++ns; }
                e4.nsrc = ns;
                e4.out_f32 = g_next;
                if (n > 0) { e4.out_split = Kbar; e4.k[0].split_scale = dt_of(n - 1) * d->b[S - 1]; }
            }
            if (launch_groupnorm_bwd_epi(sl.X, mp->norm_w[0], mp->norm_b[0], dH, 1.f, d->act, e4, need_w ? gnpart[0] : nullptr,
                                         need_w ? gnpart[1] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            ++evals;
        }
        g_cur = g_next;
    }
    if (need_w) {
        if (wgrad_finish(engine, acc1, shp, st) || wgrad_finish(engine, acc2, shp, st)) return -1;
        for (int k = 0; k < 2; ++k) {
            launch_sum_over_batch(gnpart[2 * k], grads->norm_w[k], d->batch, C, st);
            launch_sum_over_batch(gnpart[2 * k + 1], grads->norm_b[k], d->batch, C, st);
        }
    }
    return check_cuda(cudaGetLastError(), "odeblock backward (preact GN)");
}

// ---------------------------------------------------------------------------------------------
// CIFAR POST-activation right-hand side with GroupNorm (BasicBlock2, cifar10/layers.py:108-121 with the 'GN' / 'LN' /
// 'IN' normalisations of utils.py:26-36):      f(x) = act(GN2(conv2(act(GN1(conv1(x))))))
// conv1 reads split(x_i); each GroupNorm (+ activation) is its own launch; k_i = act(GN2(.)) and the Runge-Kutta stage
// combination are the epilogue of the GN2 launch.  Tape per (step, stage): P1, P2 (conv outputs), A = split(x_i),
// Hs = split(act(GN1(P1))).
// ---------------------------------------------------------------------------------------------
struct PgSlot { float *P1, *P2; __nv_bfloat16 *A, *Hs; };
static PgSlot pg_slot(void* tape, size_t E, int slot) {
    const size_t q = align_up(E * 4);
    char* p = (char*)tape + (size_t)slot * 4 * q;
    return PgSlot{(float*)p, (float*)(p + q), (__nv_bfloat16*)(p + 2 * q), (__nv_bfloat16*)(p + 3 * q)};
}

static int gn_postact_forward(const MsbOdeDesc* d, const float* x, const MsbMnistParams* mp, float* y_out, void* workspace,
                              size_t workspace_bytes, void* tape, size_t tape_bytes, cudaStream_t st) {
    const bool save = d->save_tape != 0;
    if (save && (!tape || tape_bytes < msb_odeblock_tape_bytes(d))) { set_error("tape missing or too small"); return -1; }
    if (gn_check_params(mp)) return -1;
    if (!x || !y_out || !workspace) { set_error("null pointer argument"); return -1; }
    if (workspace_bytes < msb_odeblock_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const int engine = resolve_engine(d);
    if (engine < 0) return -1;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    ConvShape shp{d->batch, d->height, d->width, C};
    Carver cv(workspace, workspace_bytes);
    void* wp[2] = {cv.take<char>(packed_w_bytes(engine, C)), cv.take<char>(packed_w_bytes(engine, C))};
    float* ybuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* kbuf[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < S - 1; ++i) kbuf[i] = cv.take<float>(E * 4);
    __nv_bfloat16* A = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* Hs = cv.take<__nv_bfloat16>(E * 4);
    float* P1 = cv.take<float>(E * 4);
    float* P2 = cv.take<float>(E * 4);
    float* xbuf = cv.take<float>(E * 4);
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    for (int k = 0; k < 2; ++k) pack_w(engine, mp->conv_w[k], wp[k], C, 0, st, d->height, d->width);
    const float* y_cur = x;
    for (int n = 0; n < N; ++n) {
        const float dt = d->time_grid[n + 1] - d->time_grid[n];
        float* y_next = (n == N - 1) ? y_out : ybuf[n & 1];
        for (int i = 0; i < S; ++i) {
            PgSlot sl = save ? pg_slot(tape, E, n * S + i) : PgSlot{P1, P2, A, Hs};
            const float* xi = (i == 0) ? y_cur : xbuf;
            launch_act_split(xi, nullptr, ACT_NONE, 1.f, sl.A, nullptr, d->batch, d->height, d->width, C, st);
            EpiParams c1 = epi_default();
            c1.out_f32 = sl.P1;
            if (run_conv(engine, sl.A, wp[0], c1, shp, st)) return -1;
            EpiParams g1 = epi_default();
            g1.act = d->act; g1.out_split = sl.Hs;
            if (launch_groupnorm_epi(sl.P1, mp->norm_w[0], mp->norm_b[0], g1, shp, mp->groups, mp->eps, st)) return -1;
            EpiParams c2 = epi_default();
            c2.out_f32 = sl.P2;
            if (run_conv(engine, sl.Hs, wp[1], c2, shp, st)) return -1;
            EpiParams g2 = epi_default();                       // k_i = act(GN2(.)) and the RK combination
            g2.act_v = d->act;
            g2.base = y_cur; g2.k[0].dt = dt;
            if (i < S - 1) {
                g2.v_out = kbuf[i];
                g2.nsrc = i;
                for (int j = 0; j < i; ++j) { g2.src[j] = kbuf[j]; g2.k[0].coef[j] = d->w[(i + 1) * MSB_MAX_STAGES + j]; }
                g2.k[0].coef_v = d->w[(i + 1) * MSB_MAX_STAGES + i];
                g2.out_f32 = xbuf;
            } else {
                g2.nsrc = S - 1;
                for (int j = 0; j < S - 1; ++j) { g2.src[j] = kbuf[j]; g2.k[0].coef[j] = d->b[j]; }
                g2.k[0].coef_v = d->b[S - 1];
                g2.out_f32 = y_next;
            }
            if (launch_groupnorm_epi(sl.P2, mp->norm_w[1], mp->norm_b[1], g2, shp, mp->groups, mp->eps, st)) return -1;
        }
        y_cur = y_next;
    }
    return check_cuda(cudaGetLastError(), "odeblock forward (postact GN)");
}

//   kbar_i --act', GN2'--> dP2 --conv2^T--> dH --act', GN1'--> dP1 --conv1^T--> xbar_i  (+ the RK adjoint combination in the
//   epilogue of the last convolution, as in the normalisation-free path)
static int gn_postact_backward(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mp, const void* tape,
                               size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads, void* workspace,
                               size_t workspace_bytes, cudaStream_t st) {
    if (gn_check_params(mp)) return -1;
    if (!grad_y || !tape || !grad_x || !workspace) { set_error("null pointer argument"); return -1; }
    if (tape_bytes < msb_odeblock_tape_bytes(d)) { set_error("tape too small"); return -1; }
    if (workspace_bytes < msb_odeblock_bwd_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const bool need_w = grads != nullptr;
    if (need_w)
        for (int i = 0; i < 2; ++i)
            if (!grads->norm_w[i] || !grads->norm_b[i] || !grads->conv_w[i]) { set_error("POSTACT_GN grads: null pointer"); return -1; }
    const int engine = resolve_engine(d);
    if (engine < 0) return -1;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    ConvShape shp{d->batch, d->height, d->width, C};
    Carver cv(workspace, workspace_bytes);
    void* wt[2] = {cv.take<char>(packed_w_bytes(engine, C)), cv.take<char>(packed_w_bytes(engine, C))};
    float* gbuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* xbar[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 1; i < S; ++i) xbar[i] = cv.take<float>(E * 4);
    __nv_bfloat16* DP2 = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* DP1 = cv.take<__nv_bfloat16>(E * 4);
    const size_t part_bytes = (size_t)wgrad_nparts(engine, shp) * 9 * C * C * 4;
    WgradAcc acc1{cv.take<float>(part_bytes), need_w ? grads->conv_w[0] : nullptr, 0, 0};
    WgradAcc acc2{cv.take<float>(part_bytes), need_w ? grads->conv_w[1] : nullptr, 0, 0};
    // one fp32 scratch tensor, reused along the chain: kbar_i (input of GN2') -> dH (conv2^T output, input of GN1') ->
    // kbar_{i-1} (written by the epilogue of conv1^T once dH has been consumed)
    float* kbar = cv.take<float>(E * 4);
    float* gnpart[4];
    for (int i = 0; i < 4; ++i) gnpart[i] = cv.take<float>((size_t)d->batch * C * 4);     // (dgamma_k, dbeta_k), k = 1, 2
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    for (int k = 0; k < 2; ++k) pack_w(engine, mp->conv_w[k], wt[k], C, 1, st, d->height, d->width);

    auto dt_of = [&](int n) { return d->time_grid[n + 1] - d->time_grid[n]; };
    int evals = 0;
    const float* g_cur = grad_y;
    for (int n = N - 1; n >= 0; --n) {
        const float dt = dt_of(n);
        float* g_next = (n == 0) ? grad_x : gbuf[n & 1];
        for (int i = S - 1; i >= 0; --i) {
            const PgSlot sl = pg_slot(const_cast<void*>(tape), E, n * S + i);
            const int acc = evals > 0;
            // dP2 = GN2'(act'(.) kbar_i);  kbar_S = dt b_S gbar (folded into the scale), kbar_i (i < S) from the previous epilogue
            const float* dy = (i == S - 1) ? g_cur : kbar;
            const float sc = (i == S - 1) ? dt * d->b[S - 1] : 1.f;
            EpiParams e = epi_default();
            e.out_split = DP2;
            if (launch_groupnorm_bwd_epi(sl.P2, mp->norm_w[1], mp->norm_b[1], dy, sc, d->act, e, need_w ? gnpart[2] : nullptr,
                                         need_w ? gnpart[3] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            // dW2 += dP2 (x) Hs_i ;  dH = dgrad_W2(dP2)
            if (need_w && run_wgrad(engine, DP2, sl.Hs, acc2, shp, st)) return -1;
            e = epi_default();
            e.out_f32 = kbar;                                         // dH (kbar_i has been consumed by GN2')
            if (run_conv(engine, DP2, wt[1], e, shp, st)) return -1;
            // dP1 = GN1'(act'(.) dH)
            e = epi_default();
            e.out_split = DP1;
            if (launch_groupnorm_bwd_epi(sl.P1, mp->norm_w[0], mp->norm_b[0], kbar, 1.f, d->act, e, need_w ? gnpart[0] : nullptr,
                                         need_w ? gnpart[1] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            // dW1 += dP1 (x) A_i ;  xbar_i = dgrad_W1(dP1) and the adjoint stage combination
            if (need_w && run_wgrad(engine, DP1, sl.A, acc1, shp, st)) return -1;
            EpiParams e4 = epi_default();
            e4.base = g_cur;
            if (i > 0) {
                e4.v_out = xbar[i];
                e4.base_is_one = 0;
                e4.k[0].base_coef = dt * d->b[i - 1];
                int ns = 0;
                for (int j = S - 1; j > i; --j) { e4.src[ns] = xbar[j]; e4.k[0].coef[ns] = d->w[j * MSB_MAX_STAGES + (i - 1)]; ++ns; }
                e4.nsrc = ns;
                e4.k[0].coef_v = d->w[i * MSB_MAX_STAGES + (i - 1)];
                e4.k[0].dt = dt;
                e4.out_f32 = kbar;                                    // = kbar_{i-1} (fp32: GN2' of the next stage reads it)
            } else {
                int ns = 0;
                for (int j = S - 1; j > 0; --j) { e4.src[ns] = xbar[j]; e4.k[0].coef[ns] = 1.f; ++ns; }
                e4.nsrc = ns;
                e4.out_f32 = g_next;
            }
            if (run_conv(engine, DP1, wt[0], e4, shp, st)) return -1;
            ++evals;
        }
        g_cur = g_next;
    }
    if (need_w) {
        if (wgrad_finish(engine, acc1, shp, st) || wgrad_finish(engine, acc2, shp, st)) return -1;
        for (int k = 0; k < 2; ++k) {
            launch_sum_over_batch(gnpart[2 * k], grads->norm_w[k], d->batch, C, st);
            launch_sum_over_batch(gnpart[2 * k + 1], grads->norm_b[k], d->batch, C, st);
        }
    }
    return check_cuda(cudaGetLastError(), "odeblock backward (postact GN)");
}

int msb_odeblock_forward(const MsbOdeDesc* d, const float* x, const float* w1, const float* w2,
                         const MsbMnistParams* mnist, float* y_out, void* workspace, size_t workspace_bytes,
                         void* tape, size_t tape_bytes, void* cuda_stream) {
    if (validate(d)) return -1;
    if (d->rhs_kind == MSB_RHS_MNIST_GN_T)
        return mnist_forward(d, x, mnist, y_out, workspace, workspace_bytes, tape, tape_bytes, (cudaStream_t)cuda_stream);
    if (d->rhs_kind == MSB_RHS_PREACT_GN)
        return gn_preact_forward(d, x, mnist, y_out, workspace, workspace_bytes, tape, tape_bytes, (cudaStream_t)cuda_stream);
    if (d->rhs_kind == MSB_RHS_POSTACT_GN)
        return gn_postact_forward(d, x, mnist, y_out, workspace, workspace_bytes, tape, tape_bytes, (cudaStream_t)cuda_stream);
    int engine = resolve_engine(d);
    if (engine < 0) return -1;
    if (!x || !w1 || !w2 || !y_out || !workspace) { set_error("null pointer argument"); return -1; }
    if (workspace_bytes < msb_odeblock_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const bool save = d->save_tape != 0;
    if (save && (!tape || tape_bytes < msb_odeblock_tape_bytes(d))) { set_error("tape missing or too small"); return -1; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);

    Carver cv(workspace, workspace_bytes);
    void* wp1 = cv.take<char>(packed_w_bytes(engine, C));
    void* wp2 = cv.take<char>(packed_w_bytes(engine, C));
    float* ybuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* kbuf[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < S - 1; ++i) kbuf[i] = cv.take<float>(E * 4);
    __nv_bfloat16* A_inf = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* Hs_inf = cv.take<__nv_bfloat16>(E * 4);
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }

    pack_w(engine, w1, wp1, C, 0, st, d->height, d->width);
    pack_w(engine, w2, wp2, C, 0, st, d->height, d->width);
    const Tabs tabs = make_tabs(d);

    auto slot_full = [&](int n, int i) {
        if (save) return tape_slot(tape, E, n * S + i);
        return TapeSlot{A_inf, nullptr, Hs_inf, nullptr};
    };
    // pre-activation RHS : f(x) = conv2(act(conv1(act(x))))   -> conv1 reads split(act(x_i)), G0 = act'(x_i)
    // post-activation RHS: f(x) = act(conv2(act(conv1(x))))   -> conv1 reads split(x_i); the G0 slot of the
    //                      tape holds G2 = act'(conv2 output), needed by the backward of k_i = act(.)
    const bool post = d->rhs_kind == MSB_RHS_POSTACT_NF;
    const int act_in = post ? ACT_NONE : d->act;
    const int MB = microbatch_images(d, tabs.K);
    const size_t img_elems = (size_t)d->height * d->width * C;
    for (int b0 = 0; b0 < d->batch; b0 += MB) {
    const size_t off = (size_t)b0 * img_elems;
    const ConvShape shp{std::min(MB, d->batch - b0), d->height, d->width, C};
    auto slot = [&](int n, int i) { return slot_at(slot_full(n, i), off); };
    {   // prologue: operand of the very first conv1
        TapeSlot s0 = slot(0, 0);
        launch_act_split(x + off, nullptr, act_in, 1.f, s0.A, post ? nullptr : s0.G0, shp.B, d->height, d->width, C, st);
    }
    const float* y_cur = x + off;
    for (int n = 0; n < N; ++n) {
        const float dt = d->time_grid[n + 1] - d->time_grid[n];      // fp32, as `t1 - t0` (rk_parametric.py:105)
        float* y_next = ((n == N - 1) ? y_out : ybuf[n & 1]) + off;
        for (int i = 0; i < S; ++i) {
            TapeSlot cur = slot(n, i);
            // conv1: P = conv(A_i, W1);  Hs_i = split(act(P)),  G1_i = act'(P)
            EpiParams e1 = epi_default();
            e1.weights_settled = 1;      // packed before the (plain) act_split launch that opens the chain: see conv_tct.cu
            e1.out_split = cur.Hs; e1.act = d->act; e1.dact_out = cur.G1;
            if (run_conv(engine, cur.A, wp1, e1, shp, st)) return -1;
            // conv2: k_i = conv(Hs_i, W2) and the Runge-Kutta combination that follows it
            EpiParams e2 = epi_default();
            e2.weights_settled = 1;
            e2.base = y_cur; e2.act = act_in; e2.slice_batch = tabs.slice_batch;
            for (int q = 0; q < tabs.K; ++q) e2.k[q].dt = dt;
            if (post) { e2.act_v = d->act; e2.dact_v_out = cur.G0; }
            // Terms whose coefficient is exactly zero for every solver slice are dropped (u = 1/2, the midpoint rule, has
            // b_1 = 0): `k_j * 0` contributes +-0 to a sum whose other terms it cannot change, so the result is the
            // reference's bit for bit (up to the sign of an exact zero) while the tensor is neither read here nor -- if no
            // later stage needs it -- written by the stage that produced it.
            auto all_zero = [&](auto coef_of) {
                for (int q = 0; q < tabs.K; ++q) if (coef_of(tabs.t[q]) != 0.f) return false;
                return true;
            };
            auto k_needed_later = [&](int j) {        // is k_j read by any stage after the one that consumes it in registers?
                for (int m = j + 2; m < S; ++m)
                    if (!all_zero([&](const MsbTableau& t) { return t.w[m * MSB_MAX_STAGES + j]; })) return true;
                return !all_zero([&](const MsbTableau& t) { return t.b[j]; });
            };
            if (i < S - 1) {
                // x_{i+1} = y + (sum_j k_j w[i+1][j]) dt          (order2stage2.py:91, order3stage3.py:100-101 ...)
                if (k_needed_later(i)) e2.v_out = kbuf[i] + off;
                int ns = 0;
                for (int j = 0; j < i; ++j) {
                    if (all_zero([&](const MsbTableau& t) { return t.w[(i + 1) * MSB_MAX_STAGES + j]; })) continue;
                    e2.src[ns] = kbuf[j] + off;
                    for (int q = 0; q < tabs.K; ++q) e2.k[q].coef[ns] = tabs.t[q].w[(i + 1) * MSB_MAX_STAGES + j];
                    ++ns;
                }
                e2.nsrc = ns;
                for (int q = 0; q < tabs.K; ++q) e2.k[q].coef_v = tabs.t[q].w[(i + 1) * MSB_MAX_STAGES + i];
                TapeSlot nx = slot(n, i + 1);
                e2.out_split = nx.A; e2.dact_out = post ? nullptr : nx.G0;
            } else {
                // y1 = y0 + (sum_j k_j b_j) dt                     (order2stage2.py:93, rk_parametric.py:106)
                int ns = 0;
                for (int j = 0; j < S - 1; ++j) {
                    if (all_zero([&](const MsbTableau& t) { return t.b[j]; })) continue;
                    e2.src[ns] = kbuf[j] + off;
                    for (int q = 0; q < tabs.K; ++q) e2.k[q].coef[ns] = tabs.t[q].b[j];
                    ++ns;
                }
                e2.nsrc = ns;
                for (int q = 0; q < tabs.K; ++q) e2.k[q].coef_v = tabs.t[q].b[S - 1];
                e2.out_f32 = y_next;
                if (n < N - 1) { TapeSlot nx = slot(n + 1, 0); e2.out_split = nx.A; e2.dact_out = post ? nullptr : nx.G0; }
            }
            if (run_conv(engine, cur.Hs, wp2, e2, shp, st)) return -1;
        }
        y_cur = y_next;
    }
    }   // micro-batches
    return check_cuda(cudaGetLastError(), "odeblock forward");
}

// grad_tab (optional, device): MSB_TABLEAU_GRAD_DOUBLES doubles, ACCUMULATED into (M = MSB_MAX_STAGES):
//   [i]             dL/db_i  = sum_n dt_n <gbar_{n+1}, k_i^n>
//   [M + i*M + j]   dL/dw_ij = sum_n dt_n <xbar_i^n, k_j^n>        (j < i)
//   [M + M*M + i]   dL/dc_i  = sum_n dt_n <kbar_i^n, df/dt(t_i, x_i)>   (time-dependent MNIST right-hand side only)
// from y1 = y + dt sum b_i k_i, x_i = y + dt sum_j w_ij k_j, t_i = t_n + c_i dt (rk_parametric_order2stage2.py:81-93 and
// analogues).  k_j is recomputed from the tape (one convolution / GroupNorm per stage).
static int odeblock_backward_impl(const MsbOdeDesc* d, const float* grad_y, const float* w1, const float* w2,
                                  const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1, float* grad_w2,
                                  double* grad_tab, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (validate(d)) return -1;
    if (d->rhs_kind == MSB_RHS_MNIST_GN_T || d->rhs_kind == MSB_RHS_PREACT_GN || d->rhs_kind == MSB_RHS_POSTACT_GN) {
        set_error("use msb_odeblock_backward_mnist for the GroupNorm right-hand sides");
        return -1;
    }
    int engine = resolve_engine(d);
    if (engine < 0) return -1;
    if (!grad_y || !w1 || !w2 || !tape || !grad_x || !workspace) { set_error("null pointer argument"); return -1; }
    if (tape_bytes < msb_odeblock_tape_bytes(d)) { set_error("tape too small"); return -1; }
    if (workspace_bytes < (grad_tab ? msb_odeblock_bwd_workspace_bytes_tableau(d) : msb_odeblock_bwd_workspace_bytes(d))) {
        set_error("workspace too small");
        return -1;
    }
    if ((grad_w1 == nullptr) != (grad_w2 == nullptr)) { set_error("grad_w1 and grad_w2 must both be given or both be NULL"); return -1; }
    const bool need_w = grad_w1 != nullptr;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    ConvShape shp{d->batch, d->height, d->width, C};

    Carver cv(workspace, workspace_bytes);
    void* wt1 = cv.take<char>(packed_w_bytes(engine, C));
    void* wt2 = cv.take<char>(packed_w_bytes(engine, C));
    float* gbuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* xbar[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};   // xbar[i], i = 1..S-1
    for (int i = 1; i < S; ++i) xbar[i] = cv.take<float>(E * 4);
    __nv_bfloat16* Kbar_full = cv.take<__nv_bfloat16>(E * 4);
    __nv_bfloat16* DP_full = cv.take<__nv_bfloat16>(E * 4);
    const size_t part_bytes = (size_t)wgrad_nparts(engine, ConvShape{d->batch, d->height, d->width, C}) * 9 * C * C * 4;
    WgradAcc acc1{cv.take<float>(part_bytes), grad_w1, 0, 0}, acc2{cv.take<float>(part_bytes), grad_w2, 0, 0};
    void* wp2f = nullptr;                                      // tableau gradients: conv2 weights packed for the forward conv
    float* kre[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    double* dot_scratch = nullptr;
    if (grad_tab) {
        wp2f = cv.take<char>(packed_w_bytes(engine, C));
        for (int i = 0; i < S; ++i) kre[i] = cv.take<float>(E * 4);
        dot_scratch = cv.take<double>(dot_scratch_bytes());
    }
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }

    pack_w(engine, w1, wt1, C, 1, st, d->height, d->width);
    pack_w(engine, w2, wt2, C, 1, st, d->height, d->width);
    if (grad_tab) pack_w(engine, w2, wp2f, C, 0, st, d->height, d->width);
    const Tabs tabs = make_tabs(d);

    auto dt_of = [&](int n) { return d->time_grid[n + 1] - d->time_grid[n]; };
    // kbar_S of the last step = dt * b_S * gbar   (post-activation RHS: times act'(conv2 output) of that stage)
    const bool post = d->rhs_kind == MSB_RHS_POSTACT_NF;
    const int MB = microbatch_images(d, tabs.K);
    const size_t img_elems = (size_t)d->height * d->width * C;
    for (int b0 = 0; b0 < d->batch; b0 += MB) {
    const size_t off = (size_t)b0 * img_elems;
    const ConvShape shp{std::min(MB, d->batch - b0), d->height, d->width, C};
    __nv_bfloat16* const Kbar = Kbar_full + 2 * off;
    __nv_bfloat16* const DP = DP_full + 2 * off;
    auto slot = [&](int n, int i) { return slot_at(tape_slot(const_cast<void*>(tape), E, n * S + i), off); };
    auto g2_of = [&](int n, int i) { return slot(n, i).G0; };
    // grad_tab[q * MSB_TABLEAU_GRAD_DOUBLES + idx] += scale <a, b> over the images of slice q (stacked solver axis: slice q
    // is the contiguous image range [q * slice_batch, (q+1) * slice_batch) and owns its own tableau, hence its own sums)
    auto dot_slices = [&](const float* a, const float* b, size_t n_mb, double scale, int idx) {
        if (tabs.K <= 1) { launch_dot_accumulate(a + off, b + off, n_mb, scale, grad_tab + idx, dot_scratch, st); return; }
        const size_t se = (size_t)tabs.slice_batch * img_elems;
        for (int q = 0; q < tabs.K; ++q)
            launch_dot_accumulate(a + q * se, b + q * se, se, scale, grad_tab + q * MSB_TABLEAU_GRAD_DOUBLES + idx, dot_scratch, st);
    };
    {
        float scales[kMaxSlices];
        for (int q = 0; q < tabs.K; ++q) scales[q] = dt_of(N - 1) * tabs.t[q].b[S - 1];
        launch_act_split_sliced(grad_y + off, post ? g2_of(N - 1, S - 1) : nullptr, ACT_NONE, scales, tabs.K, Kbar, nullptr,
                                shp.B, d->height, d->width, C, st);
    }
    const float* g_cur = grad_y + off;
    for (int n = N - 1; n >= 0; --n) {
        const float dt = dt_of(n);
        float* g_next = ((n == 0) ? grad_x : gbuf[n & 1]) + off;
        const size_t n_mb = (size_t)shp.B * img_elems;
        if (grad_tab) {
            // k_j = f(x_j) again (conv2 of the taped act(conv1(..))) and dL/db_j += dt <gbar, k_j>
            for (int j = 0; j < S; ++j) {
                EpiParams ek = epi_default();
                ek.weights_settled = 1;
                ek.v_out = kre[j] + off;
                if (post) ek.act_v = d->act;
                if (run_conv(engine, slot(n, j).Hs, wp2f, ek, shp, st)) return -1;
                dot_slices(g_cur - off, kre[j], n_mb, (double)dt, j);
            }
        }
        for (int i = S - 1; i >= 0; --i) {
            TapeSlot cur = slot(n, i);
            // Kbar = split(kbar_i).   dW2 += kbar_i (x) Hs_i
            if (need_w && run_wgrad(engine, Kbar, cur.Hs, acc2, shp, st)) return -1;
            // dP = dgrad_W2(kbar_i) * act'(P_i)
            EpiParams e3 = epi_default();
            e3.weights_settled = 1;
            e3.mul = cur.G1; e3.out_split = DP;
            if (run_conv(engine, Kbar, wt2, e3, shp, st)) return -1;
            // dW1 += dP (x) A_i
            if (need_w && run_wgrad(engine, DP, cur.A, acc1, shp, st)) return -1;
            // xbar_i = dgrad_W1(dP) * act'(x_i), then the adjoint stage combination
            EpiParams e4 = epi_default();
            e4.weights_settled = 1;
            e4.mul = post ? nullptr : cur.G0; e4.base = g_cur; e4.slice_batch = tabs.slice_batch;
            if (i > 0) {
                // kbar_{i-1} = dt b_{i-1} gbar + dt sum_{j >= i} w[j][i-1] xbar_j
                // (zero-coefficient terms are dropped as in the forward pass: b_1 = 0 for the midpoint rule means gbar is
                //  not read here)
                auto all_zero = [&](auto coef_of) {
                    for (int q = 0; q < tabs.K; ++q) if (coef_of(tabs.t[q]) != 0.f) return false;
                    return true;
                };
                e4.v_out = xbar[i] + off;
                e4.base_is_one = 0;
                if (all_zero([&](const MsbTableau& t) { return dt * t.b[i - 1]; })) e4.base = nullptr;
                int ns = 0;
                for (int j = S - 1; j > i; --j) {
                    if (all_zero([&](const MsbTableau& t) { return t.w[j * MSB_MAX_STAGES + (i - 1)]; })) continue;
                    e4.src[ns] = xbar[j] + off;
                    for (int q = 0; q < tabs.K; ++q) e4.k[q].coef[ns] = tabs.t[q].w[j * MSB_MAX_STAGES + (i - 1)];
                    ++ns;
                }
                e4.nsrc = ns;
                for (int q = 0; q < tabs.K; ++q) {
                    EpiCoef& k = e4.k[q];
                    k.base_coef = dt * tabs.t[q].b[i - 1];
                    k.coef_v = tabs.t[q].w[i * MSB_MAX_STAGES + (i - 1)];
                    k.dt = dt;
                }
                e4.out_split = Kbar;
                if (post) e4.split_mul = g2_of(n, i - 1);
            } else {
                // ybar = gbar + sum_i xbar_i ; and kbar_S of the previous step
                int ns = 0;
                for (int j = S - 1; j > 0; --j) e4.src[ns++] = xbar[j] + off;
                e4.nsrc = ns;
                for (int q = 0; q < tabs.K; ++q) e4.k[q].coef[0] = e4.k[q].coef[1] = e4.k[q].coef[2] = 1.f;
                e4.out_f32 = g_next;
                if (n > 0) {
                    e4.out_split = Kbar;
                    for (int q = 0; q < tabs.K; ++q) e4.k[q].split_scale = dt_of(n - 1) * tabs.t[q].b[S - 1];
                    if (post) e4.split_mul = g2_of(n - 1, S - 1);
                }
            }
            if (run_conv(engine, DP, wt1, e4, shp, st)) return -1;
            if (grad_tab && i > 0)                     // dL/dw_ij += dt <xbar_i, k_j>, j < i
                for (int j = 0; j < i; ++j)
                    dot_slices(xbar[i], kre[j], n_mb, (double)dt, MSB_MAX_STAGES + i * MSB_MAX_STAGES + j);
        }
        g_cur = g_next;
    }
    }   // micro-batches
    if (need_w && (wgrad_finish(engine, acc1, shp, st) || wgrad_finish(engine, acc2, shp, st))) return -1;
    return check_cuda(cudaGetLastError(), "odeblock backward");
}

int msb_odeblock_backward(const MsbOdeDesc* d, const float* grad_y, const float* w1, const float* w2,
                          const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1, float* grad_w2,
                          void* workspace, size_t workspace_bytes, void* cuda_stream) {
    return odeblock_backward_impl(d, grad_y, w1, w2, tape, tape_bytes, grad_x, grad_w1, grad_w2, nullptr, workspace,
                                  workspace_bytes, cuda_stream);
}

int msb_odeblock_backward_tableau(const MsbOdeDesc* d, const float* grad_y, const float* w1, const float* w2,
                                  const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1, float* grad_w2,
                                  double* grad_tableau, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!grad_tableau) { set_error("msb_odeblock_backward_tableau: grad_tableau is NULL"); return -1; }
    return odeblock_backward_impl(d, grad_y, w1, w2, tape, tape_bytes, grad_x, grad_w1, grad_w2, grad_tableau, workspace,
                                  workspace_bytes, cuda_stream);
}

// ---------------------------------------------------------------------------------------------
// MNIST right-hand side, backward: the adjoint of mnist_forward, stage by stage in reverse.
//   kbar_i --GN3'--> dP2 --conv2^T--> . --ReLU', GN2'--> dP1 --conv1^T--> . --ReLU', GN1'--> xbar_i
// with the Runge-Kutta adjoint combination as the epilogue of the GN1' launch, and the parameter gradients
// (x-channel weights: wgrad GEMM; time channel + bias: concat_aux_grad; gamma / beta: per-sample partials).
// ---------------------------------------------------------------------------------------------
static int odeblock_backward_mnist_impl(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mp, const void* tape,
                                        size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads, double* grad_tab,
                                        void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (validate(d)) return -1;
    if (d->rhs_kind == MSB_RHS_PREACT_GN || d->rhs_kind == MSB_RHS_POSTACT_GN) {
        if (grad_tab) { set_error("tableau gradients are not implemented for the CIFAR GroupNorm right-hand sides"); return -1; }
        return (d->rhs_kind == MSB_RHS_PREACT_GN ? gn_preact_backward : gn_postact_backward)(
            d, grad_y, mp, tape, tape_bytes, grad_x, grads, workspace, workspace_bytes, (cudaStream_t)cuda_stream);
    }
    if (d->rhs_kind != MSB_RHS_MNIST_GN_T) { set_error("msb_odeblock_backward_mnist: rhs_kind must be a GroupNorm right-hand side"); return -1; }
    if (grad_tab && d->n_solvers > 1) { set_error("tableau gradients are not implemented for a stacked solver axis"); return -1; }
    if (!grad_y || !mp || !tape || !grad_x || !workspace) { set_error("null pointer argument"); return -1; }
    for (int i = 0; i < 3; ++i) if (!mp->norm_w[i] || !mp->norm_b[i]) { set_error("MNIST params: null norm pointer"); return -1; }
    for (int i = 0; i < 2; ++i) if (!mp->conv_w[i] || !mp->conv_b[i]) { set_error("MNIST params: null conv pointer"); return -1; }
    if (tape_bytes < msb_odeblock_tape_bytes(d)) { set_error("tape too small"); return -1; }
    if (workspace_bytes < (grad_tab ? msb_odeblock_bwd_workspace_bytes_tableau(d) : msb_odeblock_bwd_workspace_bytes(d))) {
        set_error("workspace too small");
        return -1;
    }
    const bool need_w = grads != nullptr;
    if (need_w) {
        for (int i = 0; i < 3; ++i) if (!grads->norm_w[i] || !grads->norm_b[i]) { set_error("MNIST grads: null norm pointer"); return -1; }
        for (int i = 0; i < 2; ++i) if (!grads->conv_w[i] || !grads->conv_b[i]) { set_error("MNIST grads: null conv pointer"); return -1; }
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int S = d->stages, N = d->n_steps, C = d->channels;
    const size_t E = state_elems(d);
    ConvShape shp{d->batch, d->height, d->width, C};
    Carver cv(workspace, workspace_bytes);
    float* wt[2] = {cv.take<float>((size_t)9 * C * C * 4), cv.take<float>((size_t)9 * C * C * 4)};
    float* gbuf[2] = {cv.take<float>(E * 4), cv.take<float>(E * 4)};
    float* xbar[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 1; i < S; ++i) xbar[i] = cv.take<float>(E * 4);
    float* kbar = cv.take<float>(E * 4);
    float* dP = cv.take<float>(E * 4);
    float* dH = cv.take<float>(E * 4);
    __nv_bfloat16* Dsplit = cv.take<__nv_bfloat16>(E * 4);
    float* wpart = cv.take<float>((size_t)wgrad_simt_nparts(shp) * 9 * C * C * 4);
    float* gnpart[6];
    for (int i = 0; i < 6; ++i) gnpart[i] = cv.take<float>((size_t)d->batch * C * 4);     // (dgamma_k, dbeta_k), k = 1..3
    float* kre[MSB_MAX_STAGES] = {nullptr, nullptr, nullptr, nullptr};
    float* tmap[2] = {nullptr, nullptr};
    double* dot_scratch = nullptr;
    const size_t map_elems = (size_t)d->height * d->width * C;
    if (grad_tab) {
        for (int i = 0; i < S; ++i) kre[i] = cv.take<float>(E * 4);
        for (int k = 0; k < 2; ++k) tmap[k] = cv.take<float>(map_elems * 4);
        dot_scratch = cv.take<double>(dot_scratch_bytes());
    }
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    for (int k = 0; k < 2; ++k) launch_pack_w_simt(mp->conv_w[k], wt[k], C, C + 1, 1, 1, st);
    if (grad_tab)     // dP/dt of a time-concatenated convolution = its per-pixel sum of in-bounds time-channel taps
        for (int k = 0; k < 2; ++k) launch_time_tapmap(mp->conv_w[k], tmap[k], d->height, d->width, C, st);

    auto dt_of = [&](int n) { return d->time_grid[n + 1] - d->time_grid[n]; };
    int evals = 0;                                   // stage evaluations processed so far (first one overwrites the accumulators)
    const float* g_cur = grad_y;
    for (int n = N - 1; n >= 0; --n) {
        const float dt = dt_of(n);
        float* g_next = (n == 0) ? grad_x : gbuf[n & 1];
        if (grad_tab) {
            // k_j = GN3(P2_j) again, and dL/db_j += dt <gbar, k_j>
            for (int j = 0; j < S; ++j) {
                const MnistSlot sj = mnist_slot(const_cast<void*>(tape), E, n * S + j);
                EpiParams ek = epi_default();
                ek.v_out = kre[j];
                if (launch_groupnorm_epi(sj.P2, mp->norm_w[2], mp->norm_b[2], ek, shp, mp->groups, mp->eps, st)) return -1;
                launch_dot_accumulate(g_cur, kre[j], E, (double)dt, grad_tab + j, dot_scratch, st);
            }
        }
        double* const gc = grad_tab ? grad_tab + MSB_MAX_STAGES + MSB_MAX_STAGES * MSB_MAX_STAGES : nullptr;
        for (int i = S - 1; i >= 0; --i) {
            const MnistSlot sl = mnist_slot(const_cast<void*>(tape), E, n * S + i);
            const float ti = mnist_stage_time(d, n, i);
            const int acc = evals > 0;
            // kbar_S = dt b_S gbar (folded into the scale of dy); kbar_i (i < S) was formed by the previous GN1' epilogue
            const float* dy3 = (i == S - 1) ? g_cur : kbar;
            const float sc3 = (i == S - 1) ? dt * d->b[S - 1] : 1.f;
            EpiParams e = epi_default();
            e.out_f32 = dP; e.out_split = Dsplit;
            if (launch_groupnorm_bwd_epi(sl.P2, mp->norm_w[2], mp->norm_b[2], dy3, sc3, 0, e, need_w ? gnpart[4] : nullptr,
                                         need_w ? gnpart[5] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            if (gc && i > 0)      // t_i = t_n + c_i dt enters conv2 through its time channel: dL/dc_i += dt <dP2, dP2/dt>
                launch_dot_bcast_accumulate(dP, tmap[1], E, map_elems, (double)dt, gc + i, dot_scratch, st);
            if (need_w) {
                int np = 0;
                if (launch_wgrad3x3_simt(Dsplit, sl.Hs, wpart, &np, shp, st)) return -1;
                launch_wgrad_reduce(wpart, np, grads->conv_w[1], C, acc, st, C + 1, 1);
                launch_concat_aux_grad(dP, ti, grads->conv_w[1], grads->conv_b[1], acc, shp, st);
            }
            e = epi_default();
            e.out_f32 = dH;
            if (run_conv(MSB_ENGINE_SIMT, Dsplit, wt[1], e, shp, st)) return -1;
            e = epi_default();
            e.out_f32 = dP; e.out_split = Dsplit;
            if (launch_groupnorm_bwd_epi(sl.P1, mp->norm_w[1], mp->norm_b[1], dH, 1.f, ACT_RELU, e, need_w ? gnpart[2] : nullptr,
                                         need_w ? gnpart[3] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            if (gc && i > 0)      // ... and conv1 through its own
                launch_dot_bcast_accumulate(dP, tmap[0], E, map_elems, (double)dt, gc + i, dot_scratch, st);
            if (need_w) {
                int np = 0;
                if (launch_wgrad3x3_simt(Dsplit, sl.A, wpart, &np, shp, st)) return -1;
                launch_wgrad_reduce(wpart, np, grads->conv_w[0], C, acc, st, C + 1, 1);
                launch_concat_aux_grad(dP, ti, grads->conv_w[0], grads->conv_b[0], acc, shp, st);
            }
            e = epi_default();
            e.out_f32 = dH;
            if (run_conv(MSB_ENGINE_SIMT, Dsplit, wt[0], e, shp, st)) return -1;
            // xbar_i = GN1'(ReLU'(.)) and the adjoint stage combination (same algebra as the CIFAR path)
            EpiParams e4 = epi_default();
            e4.base = g_cur;
            if (i > 0) {
                e4.v_out = xbar[i];
                e4.base_is_one = 0;
                e4.k[0].base_coef = dt * d->b[i - 1];
                int ns = 0;
                for (int j = S - 1; j > i; --j) { e4.src[ns] = xbar[j]; e4.k[0].coef[ns] = d->w[j * MSB_MAX_STAGES + (i - 1)]; ++ns; }
                e4.nsrc = ns;
                e4.k[0].coef_v = d->w[i * MSB_MAX_STAGES + (i - 1)];
                e4.k[0].dt = dt;
                e4.out_f32 = kbar;
            } else {
                int ns = 0;
                for (int j = S - 1; j > 0; --j) { e4.src[ns] = xbar[j]; e4.k[0].coef[ns] = 1.f; ++ns; }
                e4.nsrc = ns;
                e4.out_f32 = g_next;
            }
            if (launch_groupnorm_bwd_epi(sl.X, mp->norm_w[0], mp->norm_b[0], dH, 1.f, ACT_RELU, e4, need_w ? gnpart[0] : nullptr,
                                         need_w ? gnpart[1] : nullptr, acc, shp, mp->groups, mp->eps, st)) return -1;
            if (grad_tab && i > 0)                     // dL/dw_ij += dt <xbar_i, k_j>, j < i
                for (int j = 0; j < i; ++j)
                    launch_dot_accumulate(xbar[i], kre[j], E, (double)dt,
                                          grad_tab + MSB_MAX_STAGES + i * MSB_MAX_STAGES + j, dot_scratch, st);
            ++evals;
        }
        g_cur = g_next;
    }
    if (need_w)
        for (int k = 0; k < 3; ++k) {
            launch_sum_over_batch(gnpart[2 * k], grads->norm_w[k], d->batch, C, st);
            launch_sum_over_batch(gnpart[2 * k + 1], grads->norm_b[k], d->batch, C, st);
        }
    return check_cuda(cudaGetLastError(), "odeblock backward (mnist)");
}

int msb_odeblock_backward_mnist(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mp, const void* tape,
                                size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads, void* workspace,
                                size_t workspace_bytes, void* cuda_stream) {
    return odeblock_backward_mnist_impl(d, grad_y, mp, tape, tape_bytes, grad_x, grads, nullptr, workspace, workspace_bytes,
                                        cuda_stream);
}

int msb_odeblock_backward_mnist_tableau(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mp, const void* tape,
                                        size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads, double* grad_tableau,
                                        void* workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!grad_tableau) { set_error("msb_odeblock_backward_mnist_tableau: grad_tableau is NULL"); return -1; }
    return odeblock_backward_mnist_impl(d, grad_y, mp, tape, tape_bytes, grad_x, grads, grad_tableau, workspace,
                                        workspace_bytes, cuda_stream);
}

int msb_act_split(const float* x, int act, void* split_out, float* dact_out, int batch, int height, int width,
                  int channels, void* cuda_stream) {
    if (!x || !split_out || channels % 4) { set_error("msb_act_split: bad arguments"); return -1; }
    launch_act_split(x, nullptr, act, 1.f, (__nv_bfloat16*)split_out, dact_out, batch, height, width, channels, (cudaStream_t)cuda_stream);
    return check_cuda(cudaGetLastError(), "act_split");
}

size_t msb_conv3x3_workspace_bytes(int channels) {
    size_t a = packed_w_bytes(MSB_ENGINE_TCGEN05, channels), b = (size_t)9 * channels * channels * 4;
    return align_up(a > b ? a : b) + 1024;
}

int msb_conv3x3(const void* split_in, const float* w_oihw, float* out, int transpose, int engine, int batch, int height,
                int width, int channels, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    MsbOdeDesc d;
    memset(&d, 0, sizeof(d));
    d.engine = engine; d.batch = batch; d.height = height; d.width = width; d.channels = channels;
    d.stages = 1; d.n_steps = 1; d.act = MSB_ACT_NONE; d.rhs_kind = MSB_RHS_PREACT_NF;
    float tg[2] = {0.f, 1.f}; d.time_grid = tg;
    if (validate(&d)) return -1;
    int eng = resolve_engine(&d);
    if (eng < 0) return -1;
    if (!split_in || !w_oihw || !out || !workspace || workspace_bytes < msb_conv3x3_workspace_bytes(channels)) {
        set_error("msb_conv3x3: bad arguments / workspace too small"); return -1;
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    pack_w(eng, w_oihw, workspace, channels, transpose, st, height, width);
    EpiParams e = epi_default();
    e.out_f32 = out;
    return run_conv(eng, (const __nv_bfloat16*)split_in, workspace, e, ConvShape{batch, height, width, channels}, st);
}

size_t msb_wgrad3x3_workspace_bytes(int channels, int engine) {
    (void)engine;
    return (size_t)160 * 9 * channels * channels * 4 + 1024;   // >= max nparts of either engine
}

int msb_wgrad3x3(const void* split_grad_out, const void* split_in, float* grad_w_oihw, int engine, int batch, int height,
                 int width, int channels, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    MsbOdeDesc d;
    memset(&d, 0, sizeof(d));
    d.engine = engine; d.batch = batch; d.height = height; d.width = width; d.channels = channels;
    d.stages = 1; d.n_steps = 1; d.act = MSB_ACT_NONE; d.rhs_kind = MSB_RHS_PREACT_NF;
    float tg[2] = {0.f, 1.f}; d.time_grid = tg;
    if (validate(&d)) return -1;
    int eng = resolve_engine(&d);
    if (eng < 0) return -1;
    ConvShape s{batch, height, width, channels};
    if (!split_grad_out || !split_in || !grad_w_oihw || !workspace ||
        workspace_bytes < (size_t)wgrad_nparts(eng, s) * 9 * channels * channels * 4) {
        set_error("msb_wgrad3x3: bad arguments / workspace too small"); return -1;
    }
    WgradAcc acc{(float*)workspace, grad_w_oihw, 0, 0};
    if (run_wgrad(eng, (const __nv_bfloat16*)split_grad_out, (const __nv_bfloat16*)split_in, acc, s, (cudaStream_t)cuda_stream))
        return -1;
    return wgrad_finish(eng, acc, s, (cudaStream_t)cuda_stream);
}

}  // extern "C"
