// Shared device-side definitions: the fused epilogue contract, activation math, bf16 hi/lo split.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace msb {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

// Bottleneck-decomposition switches (scripts/conv_bottleneck.py), compiled in only with -DMSB_CONV_DEBUG; one
// copy of the flag word per translation unit, all set by msb_debug_conv_flags().
//   1: epilogue skipped (accumulator barriers still cycle)   2: no MMAs issued   4: no weight TMA
//   8: no activation TMA   16: epilogue global stores off   32: epilogue global loads off
//   64: epilogue loads wrapped into the first 1 MB of each array (L2 hits: isolates the loads' DRAM latency)
//   128: epilogue stores wrapped into the first 1 MB of each array (isolates the stores' DRAM traffic)
#ifdef MSB_CONV_DEBUG
static __device__ int g_conv_debug = 0;
#define MSB_DBG(bit) (g_conv_debug & (bit))
#define MSB_DBG_LD(idx) ((g_conv_debug & 64) ? ((idx) & (size_t)0x3FFF8) : (idx))
#define MSB_DBG_ST(idx) ((g_conv_debug & 128) ? ((idx) & (size_t)0x3FFF8) : (idx))
#else
#define MSB_DBG(bit) 0
#define MSB_DBG_LD(idx) (idx)
#define MSB_DBG_ST(idx) (idx)
#endif

// ---------------------------------------------------------------------------------------------
// Fused epilogue applied to every accumulator element of a convolution.  One struct covers the
// forward RK stage combinations, the activation that feeds the next convolution, and the
// backward (adjoint) stage combinations -- see odeblock.cu for how each stage fills it in.
//
//   v   = acc (+ chan_bias[c] + pix_bias_scale * pix_bias[h,w,c])       bias terms: MNIST ConcatConv2d only
//   v   = act_v ? act(v) : v               (dact_v_out <- act'(pre-activation))
//   v   = mul ? v * mul[idx] : v           (backward: multiply by a saved activation derivative)
//   v_out[idx]  <- v                       (k_i in forward, xbar_i in backward)
//   s   = src0*coef0 + src1*coef1 + src2*coef2 + v*coef_v      summed left to right, no FMA,
//   out = base*base_coef + s*dt                                exactly the reference's order
//         (rk_parametric_order2stage2.py:91-93: x + k1*w21*dt ; (k1*b1 + k2*b2)*dt ; rk_parametric.py:106)
//   out_f32[idx]   <- out
//   out_split      <- hi/lo( (act ? act(out) : out) * split_scale * split_mul[idx] )   operand of the next conv
//                     (split_mul: post-activation RHS backward, kbar * act'(conv2 output))
//   dact_out[idx]  <- act'(out)
// ---------------------------------------------------------------------------------------------
// Scalar coefficients of one epilogue.  With a stacked solver axis (solver ensembling: K solvers
// integrate K copies of the batch in one launch) every slice of the batch has its own set.
constexpr int kMaxSlices = 8;
struct EpiCoef {
    float coef[3];
    float coef_v;
    float dt;
    float base_coef;
    float split_scale;
};

struct EpiParams {
    const float* mul;
    float* v_out;
    const float* base;
    const float* src[3];
    float* out_f32;
    __nv_bfloat16* out_split;
    float* dact_out;
    float* dact_v_out;
    const float* chan_bias;
    const float* pix_bias;
    const float* split_mul;
    EpiCoef k[kMaxSlices];     // k[0] unless slice_batch > 0: then slice = image index / slice_batch
    float pix_bias_scale;
    int slice_batch;
    int nsrc;
    int act;
    int act_v;
    int base_is_one;   // base_coef == 1 -> skip the multiply (keeps forward bit-order exact)
    int weights_settled;   // the packed weights were written by a fully ordered (non-programmatic) launch that is NOT this
                           // kernel's immediate predecessor: a programmatically launched kernel may read them before its
                           // griddepcontrol.wait (set by the ODE-block loops; 0 = read them after the wait)
};

__host__ inline EpiParams epi_default() {
    EpiParams e;
    e.mul = nullptr; e.v_out = nullptr; e.base = nullptr;
    e.src[0] = e.src[1] = e.src[2] = nullptr;
    e.out_f32 = nullptr; e.out_split = nullptr; e.dact_out = nullptr; e.dact_v_out = nullptr;
    e.chan_bias = nullptr; e.pix_bias = nullptr; e.split_mul = nullptr;
    for (int i = 0; i < kMaxSlices; ++i) {
        e.k[i].coef[0] = e.k[i].coef[1] = e.k[i].coef[2] = 0.f;
        e.k[i].base_coef = 1.f; e.k[i].coef_v = 1.f; e.k[i].dt = 1.f; e.k[i].split_scale = 1.f;
    }
    e.pix_bias_scale = 0.f; e.slice_batch = 0;
    e.nsrc = 0; e.act = ACT_NONE; e.act_v = ACT_NONE; e.base_is_one = 1; e.weights_settled = 0;
    return e;
}

// The fp32 NHWC arrays an epilogue READS (all indexed like the output): candidates for an L2 prefetch.
// Slot i of the fixed list (null = not read); callers unroll over kEpiLoadSlots.
constexpr int kEpiLoadSlots = 6;
__device__ __forceinline__ const float* epi_load_operand(const EpiParams& e, int i) {
    switch (i) {
        case 0: return e.mul;
        case 1: return e.base;
        case 2: return e.nsrc > 0 ? e.src[0] : nullptr;
        case 3: return e.nsrc > 1 ? e.src[1] : nullptr;
        case 4: return e.nsrc > 2 ? e.src[2] : nullptr;
        default: return e.split_mul;
    }
}

__device__ __forceinline__ const EpiCoef& epi_coef(const EpiParams& e, int n) {
    return e.k[e.slice_batch > 0 ? n / e.slice_batch : 0];
}

// ---- activations -----------------------------------------------------------------------------
// GeLU (exact-erf form, cifar10/utils.py:67-68) and its derivative from ONE exponential:
//   Phi(x) = 1 - h (x >= 0),  h (x < 0),   h = 0.5 * erfc(|x| / sqrt 2) = Q(t) * E,
//   E = exp(-x^2 / 2),  t = 1 / (1 + p |x| / sqrt 2),  Q = degree-7 polynomial in t without constant term
// (same shape as Abramowitz-Stegun 7.1.26 but refitted: p = 0.45, minimax on the absolute error of erf,
// approximation error 1e-9; evaluated in fp32 the max abs error is 1.8e-7 on Phi and 1.9e-7 on gelu',
// checked offline against fp64 over [-13, 13]).  E is also the Gaussian of the derivative
// gelu'(x) = Phi(x) + x * E / sqrt(2 pi).  Branch-free: 2 MUFU + ~17 FMA-pipe instructions per element.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_both(float x, float& a, float& d) {
    const float ax = fabsf(x);
    const float E = ex2_approx((x * x) * -0.72134752044448170368f);        // exp(-x^2/2)
    const float t = rcp_approx(fmaf(ax, 0.31819805153394638f, 1.0f));      // 0.45 / sqrt 2
    float q = 0.04628484081253973f;                                        // 0.5 * c7
    q = fmaf(q, t, -0.2470522577736345f);
    q = fmaf(q, t, 0.4285840718840338f);
    q = fmaf(q, t, -0.18790157886011314f);
    q = fmaf(q, t, 0.23008541295425833f);
    q = fmaf(q, t, 0.1005046642482306f);
    q = fmaf(q, t, 0.12949484725821528f);
    const float h = (q * t) * E;
    const float cdf = x >= 0.f ? 1.0f - h : h;
    a = x * cdf;
    d = fmaf(x, E * 0.39894228040143267794f, cdf);
}
template <int ACT>
__device__ __forceinline__ void act_both_t(float x, float& a, float& d) {
    if (ACT == 1) {                      // ACT_GELU
        gelu_both(x, a, d);
    } else if (ACT == 2) {               // ACT_RELU
        a = x > 0.f ? x : 0.f;
        d = x > 0.f ? 1.f : 0.f;
    } else {
        a = x; d = 1.f;
    }
}

__device__ __forceinline__ float act_f(int act, float x) {
    float a, d;
    if (act == ACT_GELU) { act_both_t<1>(x, a, d); return a; }
    if (act == ACT_RELU) return x > 0.f ? x : 0.f;
    return x;
}
__device__ __forceinline__ float dact_f(int act, float x) {
    float a, d;
    if (act == ACT_GELU) { act_both_t<1>(x, a, d); return d; }
    if (act == ACT_RELU) return x > 0.f ? 1.f : 0.f;
    return 1.f;
}
// value and derivative together (shares the exponential)
__device__ __forceinline__ void act_both(int act, float x, float& a, float& d) {
    if (act == ACT_GELU) act_both_t<1>(x, a, d);
    else if (act == ACT_RELU) act_both_t<2>(x, a, d);
    else act_both_t<0>(x, a, d);
}

// ---- bf16 hi/lo split --------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// split tensor index: [B][H][2][W][C]
__device__ __forceinline__ size_t split_index(int n, int h, int plane, int w, int c, int H, int W, int C) {
    return ((((size_t)n * H + h) * 2 + plane) * W + w) * (size_t)C + c;
}

// ---- the epilogue itself (scalar form; engines may vectorise around it) ------------------------
// idx = NHWC linear index of the element, (n,h,w,c) its coordinates.
__device__ __forceinline__ void epilogue_apply(const EpiParams& e, float acc, size_t idx,
                                               int n, int h, int w, int c, int H, int W, int C) {
    const EpiCoef& k = epi_coef(e, n);
    float v = acc;
    if (e.chan_bias) v = __fadd_rn(v, e.chan_bias[c]);
    if (e.pix_bias) v = __fadd_rn(v, __fmul_rn(e.pix_bias_scale, e.pix_bias[((size_t)h * W + w) * C + c]));
    if (e.act_v != ACT_NONE) {
        float a, d;
        act_both(e.act_v, v, a, d);
        if (e.dact_v_out) e.dact_v_out[idx] = d;
        v = a;
    }
    if (e.mul) v = __fmul_rn(v, e.mul[idx]);
    if (e.v_out) e.v_out[idx] = v;
    float s;
    if (e.nsrc == 0) {
        s = __fmul_rn(v, k.coef_v);
    } else {
        s = __fmul_rn(e.src[0][idx], k.coef[0]);
        if (e.nsrc > 1) s = __fadd_rn(s, __fmul_rn(e.src[1][idx], k.coef[1]));
        if (e.nsrc > 2) s = __fadd_rn(s, __fmul_rn(e.src[2][idx], k.coef[2]));
        s = __fadd_rn(s, __fmul_rn(v, k.coef_v));
    }
    float out = __fmul_rn(s, k.dt);
    if (e.base) {
        float bv = e.base[idx];
        if (!e.base_is_one) bv = __fmul_rn(bv, k.base_coef);
        out = __fadd_rn(bv, out);
    }
    if (e.out_f32) e.out_f32[idx] = out;
    if (e.out_split || e.dact_out) {
        float a, d;
        act_both(e.act, out, a, d);
        if (e.dact_out) e.dact_out[idx] = d;
        if (e.out_split) {
            if (e.split_mul) a = __fmul_rn(a, e.split_mul[idx]);
            a = __fmul_rn(a, k.split_scale);
            __nv_bfloat16 hi, lo;
            split_bf16(a, hi, lo);
            e.out_split[split_index(n, h, 0, w, c, H, W, C)] = hi;
            e.out_split[split_index(n, h, 1, w, c, H, W, C)] = lo;
        }
    }
}


// ---- pipelined form of the same epilogue (tensor-core engine) --------------------------------------
// A thread owns N consecutive pixels of ONE channel (element j at idx0 + j*stride).  The operand
// loads (prefetch) are issued a whole chunk ahead of their use so their HBM/L2 latency overlaps the
// MMA of the tile and the math of the previous chunk; epi_finish() is the arithmetic of
// epilogue_apply() on N elements.  (Bias terms are SIMT-engine only.)
template <int N> struct EpiOperands {
    float mul[N];
    float base[N];
    float src[3][N];
    float smul[N];
};

__device__ __forceinline__ float ldg_stream(const float* p) {
    if (MSB_DBG(32)) return 1.f;
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int N>
__device__ __forceinline__ void epi_prefetch(const EpiParams& e, size_t idx0, int stride, EpiOperands<N>& r) {
    if (e.mul) {
#pragma unroll
        for (int j = 0; j < N; ++j) r.mul[j] = ldg_stream(e.mul + idx0 + (size_t)j * stride);
    }
    if (e.base) {
#pragma unroll
        for (int j = 0; j < N; ++j) r.base[j] = ldg_stream(e.base + idx0 + (size_t)j * stride);
    }
    if (e.nsrc > 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) r.src[0][j] = ldg_stream(e.src[0] + idx0 + (size_t)j * stride);
    }
    if (e.nsrc > 1) {
#pragma unroll
        for (int j = 0; j < N; ++j) r.src[1][j] = ldg_stream(e.src[1] + idx0 + (size_t)j * stride);
    }
    if (e.nsrc > 2) {
#pragma unroll
        for (int j = 0; j < N; ++j) r.src[2][j] = ldg_stream(e.src[2] + idx0 + (size_t)j * stride);
    }
    if (e.split_mul) {
#pragma unroll
        for (int j = 0; j < N; ++j) r.smul[j] = ldg_stream(e.split_mul + idx0 + (size_t)j * stride);
    }
}

// split_idx0 = index of element 0 in the hi plane of out_split; lo plane is plane_stride further.
// ACT = the activation `e.act` as a compile-time constant.  Every phase is a fully unrolled,
// straight-line loop over the N independent elements (uniform flags are tested outside the loops).
template <int N, int ACT>
__device__ __forceinline__ void epi_finish(const EpiParams& e_in, const EpiCoef& k, const float* acc, const EpiOperands<N>& r,
                                           size_t idx0, int stride, size_t split_idx0, size_t plane_stride) {
#ifdef MSB_CONV_DEBUG
    EpiParams e = e_in;
    if (MSB_DBG(16)) {       // stores off: keep the math alive through one predicated-off path
        const bool keep = acc[0] == 123.456f;
        if (!keep) { e.v_out = nullptr; e.out_f32 = nullptr; e.dact_out = nullptr; e.dact_v_out = nullptr; e.out_split = nullptr; }
    }
#else
    const EpiParams& e = e_in;
#endif
    float v[N], o[N];
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = acc[j];
    if (e.act_v != ACT_NONE) {            // post-activation RHS only (rare): generic path
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float a, d;
            act_both(e.act_v, v[j], a, d);
            if (e.dact_v_out) e.dact_v_out[idx0 + (size_t)j * stride] = d;
            v[j] = a;
        }
    }
    if (e.mul) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = __fmul_rn(v[j], r.mul[j]);
    }
    if (e.v_out) {
#pragma unroll
        for (int j = 0; j < N; ++j) e.v_out[idx0 + (size_t)j * stride] = v[j];
    }
    if (e.nsrc == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = __fmul_rn(v[j], k.coef_v);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = __fmul_rn(r.src[0][j], k.coef[0]);
        if (e.nsrc > 1) {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(r.src[1][j], k.coef[1]));
        }
        if (e.nsrc > 2) {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(r.src[2][j], k.coef[2]));
        }
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(v[j], k.coef_v));
    }
#pragma unroll
    for (int j = 0; j < N; ++j) o[j] = __fmul_rn(o[j], k.dt);
    if (e.base) {
        if (e.base_is_one) {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(r.base[j], o[j]);
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(__fmul_rn(r.base[j], k.base_coef), o[j]);
        }
    }
    if (e.out_f32) {
#pragma unroll
        for (int j = 0; j < N; ++j) e.out_f32[idx0 + (size_t)j * stride] = o[j];
    }
    if (e.out_split || e.dact_out) {
        float a[N], d[N];
#pragma unroll
        for (int j = 0; j < N; ++j) act_both_t<ACT>(o[j], a[j], d[j]);
        if (e.dact_out) {
#pragma unroll
            for (int j = 0; j < N; ++j) e.dact_out[idx0 + (size_t)j * stride] = d[j];
        }
        if (e.out_split) {
            if (e.split_mul) {
#pragma unroll
                for (int j = 0; j < N; ++j) a[j] = __fmul_rn(a[j], r.smul[j]);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) {
                __nv_bfloat16 hi, lo;
                split_bf16(__fmul_rn(a[j], k.split_scale), hi, lo);
                e.out_split[split_idx0 + (size_t)j * stride] = hi;
                e.out_split[split_idx0 + plane_stride + (size_t)j * stride] = lo;
            }
        }
    }
}

}  // namespace msb

// ---- channel-vector form of the same epilogue (pixel-major tensor-core engine, conv_tcp.cu) ------------
// A thread owns 8 CONSECUTIVE CHANNELS of one pixel: element j at idx0 + j.  Global accesses are whole
// 32-byte sectors (256-bit loads / stores, sm_100+), bf16 hi/lo planes get one 16-byte store each.
// Arithmetic and evaluation order are those of epilogue_apply() / epi_finish().
namespace msb {

// MAXSRC / SMUL: how many src[] arrays and whether split_mul can occur (compile-time bounds on the run-time
// e.nsrc / e.split_mul: a kernel instantiated for the common RK2 / Euler launches holds 24 operand registers instead
// of 48 -- the callers pick the instantiation from the EpiParams they launch with, see epi_is_lean()).
template <int MAXSRC_, bool SMUL_> struct EpiVec8T {
    static constexpr int MAXSRC = MAXSRC_;
    static constexpr bool SMUL = SMUL_;
    float mul[8];
    float base[8];
    float src[MAXSRC_ > 0 ? MAXSRC_ : 1][8];
    float smul[SMUL_ ? 8 : 1];
};
typedef EpiVec8T<3, true> EpiVec8;
__host__ inline bool epi_is_lean(const EpiParams& e) { return e.nsrc <= 1 && e.split_mul == nullptr; }

__device__ __forceinline__ void ldg256_stream(const float* p, float* v) {
#ifdef MSB_CONV_DEBUG
    if (g_conv_debug & 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 1.f;
        return;
    }
#endif
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
#ifdef MSB_CONV_DEBUG
    if ((g_conv_debug & 16) && v[0] != 123.456f) return;
#endif
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

template <class OPS>
__device__ __forceinline__ void epi_prefetch_vec8(const EpiParams& e, size_t idx0_in, OPS& r) {
    const size_t idx0 = MSB_DBG_LD(idx0_in);
    if (e.mul) ldg256_stream(e.mul + idx0, r.mul);
    if (e.base) ldg256_stream(e.base + idx0, r.base);
    if (OPS::MAXSRC > 0 && e.nsrc > 0) ldg256_stream(e.src[0] + idx0, r.src[0]);
    if (OPS::MAXSRC > 1 && e.nsrc > 1) ldg256_stream(e.src[1] + idx0, r.src[1]);
    if (OPS::MAXSRC > 2 && e.nsrc > 2) ldg256_stream(e.src[2] + idx0, r.src[2]);
    if (OPS::SMUL && e.split_mul) ldg256_stream(e.split_mul + idx0, r.smul);
}

// ---- packed fp32 arithmetic (sm_100: mul / add / sub / fma .f32x2, two IEEE operations per issued instruction) ----
// The vector epilogue is issue-bound (ncu: the epilogue warps keep the schedulers ~55 % busy and 36 % of their
// instructions are FMUL / FFMA / FADD), so the same round-to-nearest operations are issued on PAIRS of adjacent
// channels.  Every lane of a packed operation is the IEEE operation of the scalar code (explicit .rn, no contraction):
// results are bit-identical to epilogue_apply() / epi_finish().
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2_t bc2(float c) { return pk2(c, c); }

// gelu_both() on a pair (same operations in the same order, lane by lane)
__device__ __forceinline__ void gelu_both2(f32x2_t x, f32x2_t& a, f32x2_t& d) {
    float x0, x1;
    upk2(x, x0, x1);
    const f32x2_t ax = x & 0x7fffffff7fffffffull;
    float e0, e1;
    upk2(mul2(mul2(x, x), bc2(-0.72134752044448170368f)), e0, e1);
    const f32x2_t E = pk2(ex2_approx(e0), ex2_approx(e1));
    float t0, t1;
    upk2(fma2(ax, bc2(0.31819805153394638f), bc2(1.0f)), t0, t1);
    const f32x2_t t = pk2(rcp_approx(t0), rcp_approx(t1));
    f32x2_t q = bc2(0.04628484081253973f);
    q = fma2(q, t, bc2(-0.2470522577736345f));
    q = fma2(q, t, bc2(0.4285840718840338f));
    q = fma2(q, t, bc2(-0.18790157886011314f));
    q = fma2(q, t, bc2(0.23008541295425833f));
    q = fma2(q, t, bc2(0.1005046642482306f));
    q = fma2(q, t, bc2(0.12949484725821528f));
    const f32x2_t h = mul2(mul2(q, t), E);
    const f32x2_t omh = sub2(bc2(1.0f), h);
    float h0, h1, m0, m1;
    upk2(h, h0, h1);
    upk2(omh, m0, m1);
    const f32x2_t cdf = pk2(x0 >= 0.f ? m0 : h0, x1 >= 0.f ? m1 : h1);
    a = mul2(x, cdf);
    d = fma2(x, mul2(E, bc2(0.39894228040143267794f)), cdf);
}
template <int ACT>
__device__ __forceinline__ void act_both2_t(f32x2_t x, f32x2_t& a, f32x2_t& d) {
    if (ACT == 1) {
        gelu_both2(x, a, d);
    } else if (ACT == 2) {
        float x0, x1;
        upk2(x, x0, x1);
        a = pk2(x0 > 0.f ? x0 : 0.f, x1 > 0.f ? x1 : 0.f);
        d = pk2(x0 > 0.f ? 1.f : 0.f, x1 > 0.f ? 1.f : 0.f);
    } else {
        a = x; d = bc2(1.f);
    }
}
__device__ __forceinline__ void stg256_2(float* p, const f32x2_t* v) {
#ifdef MSB_CONV_DEBUG
    if ((g_conv_debug & 16) && v[0] != 0x12345678ull) return;
#endif
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
}

// The packed form of epi_finish_vec8 (below): same contract, same bits.
// `next` is called as soon as the prefetched operands `r` have been consumed -- before the activation / split arithmetic and
// the stores -- so that the caller can issue the operand loads of its NEXT element group into the same registers and have
// their latency covered by this group's arithmetic (they are asm volatile, like the stores: the compiler never hoists them).
struct EpiNoNext { __device__ __forceinline__ void operator()() const {} };
template <int ACT, class OPS, class NEXT = EpiNoNext>
__device__ __forceinline__ void epi_finish_vec8_x2(const EpiParams& e, const EpiCoef& k, const float* acc, const OPS& r,
                                                   size_t idx0_in, size_t split_idx0_in, size_t plane_stride, NEXT next = NEXT()) {
    constexpr int P = 4;
    const size_t idx0 = MSB_DBG_ST(idx0_in), split_idx0 = MSB_DBG_ST(split_idx0_in);
    f32x2_t v[P], o[P];
#pragma unroll
    for (int j = 0; j < P; ++j) v[j] = pk2(acc[2 * j], acc[2 * j + 1]);
    if (e.act_v != ACT_NONE) {            // post-activation RHS only (rare): generic scalar path
        float av[8], d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) act_both(e.act_v, acc[j], av[j], d[j]);
#pragma unroll
        for (int j = 0; j < P; ++j) v[j] = pk2(av[2 * j], av[2 * j + 1]);
        if (e.dact_v_out) stg256(e.dact_v_out + idx0, d);
    }
    if (e.mul) {
#pragma unroll
        for (int j = 0; j < P; ++j) v[j] = mul2(v[j], pk2(r.mul[2 * j], r.mul[2 * j + 1]));
    }
    if (e.v_out) stg256_2(e.v_out + idx0, v);
    const f32x2_t cv = bc2(k.coef_v);
    if (e.nsrc == 0) {
#pragma unroll
        for (int j = 0; j < P; ++j) o[j] = mul2(v[j], cv);
    } else {
        const f32x2_t c0 = bc2(k.coef[0]);
#pragma unroll
        for (int j = 0; j < P; ++j) o[j] = mul2(pk2(r.src[0][2 * j], r.src[0][2 * j + 1]), c0);
        if (OPS::MAXSRC > 1 && e.nsrc > 1) {
            const f32x2_t c1 = bc2(k.coef[1]);
#pragma unroll
            for (int j = 0; j < P; ++j) o[j] = add2(o[j], mul2(pk2(r.src[OPS::MAXSRC > 1 ? 1 : 0][2 * j], r.src[OPS::MAXSRC > 1 ? 1 : 0][2 * j + 1]), c1));
        }
        if (OPS::MAXSRC > 2 && e.nsrc > 2) {
            const f32x2_t c2 = bc2(k.coef[2]);
#pragma unroll
            for (int j = 0; j < P; ++j) o[j] = add2(o[j], mul2(pk2(r.src[OPS::MAXSRC > 2 ? 2 : 0][2 * j], r.src[OPS::MAXSRC > 2 ? 2 : 0][2 * j + 1]), c2));
        }
#pragma unroll
        for (int j = 0; j < P; ++j) o[j] = add2(o[j], mul2(v[j], cv));
    }
    const f32x2_t dt2 = bc2(k.dt);
#pragma unroll
    for (int j = 0; j < P; ++j) o[j] = mul2(o[j], dt2);
    if (e.base) {
        if (e.base_is_one) {
#pragma unroll
            for (int j = 0; j < P; ++j) o[j] = add2(pk2(r.base[2 * j], r.base[2 * j + 1]), o[j]);
        } else {
            const f32x2_t bcf = bc2(k.base_coef);
#pragma unroll
            for (int j = 0; j < P; ++j) o[j] = add2(mul2(pk2(r.base[2 * j], r.base[2 * j + 1]), bcf), o[j]);
        }
    }
    const bool r_used_late = OPS::SMUL && e.split_mul != nullptr;     // post-activation backward only
    if (!r_used_late) next();
    if (e.out_f32) stg256_2(e.out_f32 + idx0, o);
    if (e.out_split || e.dact_out) {
        f32x2_t a[P], d[P];
#pragma unroll
        for (int j = 0; j < P; ++j) act_both2_t<ACT>(o[j], a[j], d[j]);
        if (e.dact_out) stg256_2(e.dact_out + idx0, d);
        if (e.out_split) {
            if (OPS::SMUL && e.split_mul) {
#pragma unroll
                for (int j = 0; j < P; ++j) a[j] = mul2(a[j], pk2(r.smul[OPS::SMUL ? 2 * j : 0], r.smul[OPS::SMUL ? 2 * j + 1 : 0]));
            }
            const f32x2_t ss = bc2(k.split_scale);
            uint32_t hi[P], lo[P];
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const f32x2_t x = mul2(a[j], ss);
                float x0, x1;
                upk2(x, x0, x1);
                uint32_t h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));     // element 2j in the low half
                const f32x2_t hf = pk2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u));
                float l0, l1;
                upk2(sub2(x, hf), l0, l1);
                uint32_t l;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(l1), "f"(l0));
                hi[j] = h; lo[j] = l;
            }
#ifdef MSB_CONV_DEBUG
            if ((g_conv_debug & 16) && hi[0] != 0x12345678u) return;
#endif
            *reinterpret_cast<uint4*>(e.out_split + split_idx0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(e.out_split + split_idx0 + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (r_used_late) next();
}

// split_idx0 = index of element 0 in the hi plane of out_split; the lo plane is plane_stride further.
template <int ACT>
__device__ __forceinline__ void epi_finish_vec8(const EpiParams& e, const EpiCoef& k, const float* acc, const EpiVec8& r,
                                                size_t idx0_in, size_t split_idx0_in, size_t plane_stride) {
    constexpr int N = 8;
    const size_t idx0 = MSB_DBG_ST(idx0_in), split_idx0 = MSB_DBG_ST(split_idx0_in);
    float v[N], o[N];
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = acc[j];
    if (e.act_v != ACT_NONE) {            // post-activation RHS only (rare): generic path
        float d[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float a;
            act_both(e.act_v, v[j], a, d[j]);
            v[j] = a;
        }
        if (e.dact_v_out) stg256(e.dact_v_out + idx0, d);
    }
    if (e.mul) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = __fmul_rn(v[j], r.mul[j]);
    }
    if (e.v_out) stg256(e.v_out + idx0, v);
    if (e.nsrc == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = __fmul_rn(v[j], k.coef_v);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = __fmul_rn(r.src[0][j], k.coef[0]);
        if (e.nsrc > 1) {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(r.src[1][j], k.coef[1]));
        }
        if (e.nsrc > 2) {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(r.src[2][j], k.coef[2]));
        }
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(v[j], k.coef_v));
    }
#pragma unroll
    for (int j = 0; j < N; ++j) o[j] = __fmul_rn(o[j], k.dt);
    if (e.base) {
        if (e.base_is_one) {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(r.base[j], o[j]);
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) o[j] = __fadd_rn(__fmul_rn(r.base[j], k.base_coef), o[j]);
        }
    }
    if (e.out_f32) stg256(e.out_f32 + idx0, o);
    if (e.out_split || e.dact_out) {
        float a[N], d[N];
#pragma unroll
        for (int j = 0; j < N; ++j) act_both_t<ACT>(o[j], a[j], d[j]);
        if (e.dact_out) stg256(e.dact_out + idx0, d);
        if (e.out_split) {
            if (e.split_mul) {
#pragma unroll
                for (int j = 0; j < N; ++j) a[j] = __fmul_rn(a[j], r.smul[j]);
            }
            uint32_t hi[N / 2], lo[N / 2];
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                const float x0 = __fmul_rn(a[j], k.split_scale), x1 = __fmul_rn(a[j + 1], k.split_scale);
                // packed round-to-nearest bf16 pair (element j in the low half = lower address)
                uint32_t h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
                const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
                uint32_t l;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
                hi[j / 2] = h; lo[j / 2] = l;
            }
#ifdef MSB_CONV_DEBUG
            if ((g_conv_debug & 16) && hi[0] != 0x12345678u) return;
#endif
            *reinterpret_cast<uint4*>(e.out_split + split_idx0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(e.out_split + split_idx0 + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// which implementation the tensor-core engines call (MSB_EPI_X2=0 builds the scalar one for A/B runs)
#ifndef MSB_EPI_X2
#define MSB_EPI_X2 1
#endif
template <int ACT, class OPS, class NEXT = EpiNoNext>
__device__ __forceinline__ void epi_finish_v8(const EpiParams& e, const EpiCoef& k, const float* acc, const OPS& r,
                                              size_t idx0, size_t split_idx0, size_t plane_stride, NEXT next = NEXT()) {
#if MSB_EPI_X2
    epi_finish_vec8_x2<ACT, OPS, NEXT>(e, k, acc, r, idx0, split_idx0, plane_stride, next);
#else
    static_assert(OPS::MAXSRC == 3 && OPS::SMUL, "the scalar A/B build has no lean instantiation");
    epi_finish_vec8<ACT>(e, k, acc, r, idx0, split_idx0, plane_stride);
    next();
#endif
}

}  // namespace msb
