// SIMT engine: plain fp32 FFMA implicit-GEMM 3x3 convolution and weight-gradient kernels.
//
// Purpose: (1) shapes the tcgen05 engine does not cover (small / odd images, the 6x6 MNIST state),
// (2) an independent on-device check of the tcgen05 engine with identical operands and epilogue.
// Operands are the same bf16 hi/lo split tensors the tensor-core engine reads; here hi+lo is
// rebuilt in fp32 and multiplied on the CUDA cores.  Smem-tiled 64x64 output tiles, 4x4 per thread.
#include "msb_internal.h"

namespace msb {

namespace {
constexpr int TP = 64;   // pixels per block tile
constexpr int TC = 64;   // output channels per block tile
constexpr int TK = 16;   // input channels per k-slice
}

__global__ void __launch_bounds__(256) conv3x3_simt_kernel(const __nv_bfloat16* __restrict__ in,
                                                           const float* __restrict__ wp, EpiParams epi,
                                                           int B, int H, int W, int C) {
    __shared__ float s_in[TK][TP + 4];
    __shared__ float s_w[TK][TC];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;           // tx: 4 output channels, ty: 4 pixels
    const long long P = (long long)B * H * W;
    const long long p0 = (long long)blockIdx.x * TP;
    const int co0 = blockIdx.y * TC;

    // loader role: pixel lp, 4 consecutive input channels starting at lk
    const int lp = tid >> 2, lk = (tid & 3) * 4;
    long long lpix = p0 + lp;
    int ln = 0, lh = 0, lw = 0;
    bool lvalid = lpix < P;
    if (lvalid) { lw = (int)(lpix % W); long long t = lpix / W; lh = (int)(t % H); ln = (int)(t / H); }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        const int ih = lh + dh, iw = lw + dw;
        const bool inb = lvalid && ih >= 0 && ih < H && iw >= 0 && iw < W;
        for (int k0 = 0; k0 < C; k0 += TK) {
            // ---- stage input slice (hi + lo) ----
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (inb && k0 + lk < C) {
                size_t o_hi = split_index(ln, ih, 0, iw, k0 + lk, H, W, C);
                size_t o_lo = split_index(ln, ih, 1, iw, k0 + lk, H, W, C);
                uint2 rh = *reinterpret_cast<const uint2*>(in + o_hi);
                uint2 rl = *reinterpret_cast<const uint2*>(in + o_lo);
                const __nv_bfloat16* ph = reinterpret_cast<const __nv_bfloat16*>(&rh);
                const __nv_bfloat16* pl = reinterpret_cast<const __nv_bfloat16*>(&rl);
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __bfloat162float(ph[q]) + __bfloat162float(pl[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) s_in[lk + q][lp] = v[q];
            // ---- stage weight slice [TK][TC] ----
            {
                int kk = tid >> 4, cc = (tid & 15) * 4;
                float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + kk < C && co0 + cc < C)
                    wv = *reinterpret_cast<const float4*>(wp + ((size_t)tap * C + (k0 + kk)) * C + co0 + cc);
                *reinterpret_cast<float4*>(&s_w[kk][cc]) = wv;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                float a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = s_in[k][ty + 16 * i];
                float4 b = *reinterpret_cast<const float4*>(&s_w[k][tx * 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                    acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                    acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                    acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long pix = p0 + ty + 16 * i;
        if (pix >= P) continue;
        int w = (int)(pix % W); long long t = pix / W; int h = (int)(t % H); int n = (int)(t / H);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = co0 + tx * 4 + j;
            if (c < C) epilogue_apply(epi, acc[i][j], (size_t)pix * C + c, n, h, w, c, H, W, C);
        }
    }
}

int launch_conv3x3_simt(const __nv_bfloat16* split_in, const float* w_packed, const EpiParams& epi,
                        ConvShape s, cudaStream_t st) {
    if (s.C % 4 != 0) { set_error("SIMT conv: channels must be a multiple of 4 (got %d)", s.C); return -1; }
    long long P = (long long)s.B * s.H * s.W;
    dim3 grid((unsigned)((P + TP - 1) / TP), (unsigned)((s.C + TC - 1) / TC));
    conv3x3_simt_kernel<<<grid, 256, 0, st>>>(split_in, w_packed, epi, s.B, s.H, s.W, s.C);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv3x3_simt launch");
}

// ---------------------------------------------------------------------------------------------
// wgrad: partial[part][tap][ci][co] = sum over the part's pixels of in[p + tap][ci] * gout[p][co]
// grid = (taps * ci-tiles * co-tiles, nparts); block tile 64 ci x 64 co, 4x4 per thread,
// pixels consumed 16 at a time through smem.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad3x3_simt_kernel(const __nv_bfloat16* __restrict__ gout,
                                                            const __nv_bfloat16* __restrict__ in,
                                                            float* __restrict__ partial, int B, int H, int W, int C,
                                                            long long pix_per_part) {
    __shared__ float s_i[TK][TC + 4];   // [pixel][ci]
    __shared__ float s_g[TK][TC];       // [pixel][co]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;   // tx -> 4 co, ty -> 4 ci
    const int ctiles = (C + TC - 1) / TC;
    int bx = blockIdx.x;
    const int cot = bx % ctiles; bx /= ctiles;
    const int cit = bx % ctiles; bx /= ctiles;
    const int tap = bx;
    const int dh = tap / 3 - 1, dw = tap % 3 - 1;
    const long long P = (long long)B * H * W;
    const long long pbeg = (long long)blockIdx.y * pix_per_part;
    const long long pend = pbeg + pix_per_part < P ? pbeg + pix_per_part : P;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lp = tid >> 4, lc = (tid & 15) * 4;   // loader: pixel lp of the slice, 4 channels at lc
    for (long long q0 = pbeg; q0 < pend; q0 += TK) {
        long long pix = q0 + lp;
        float vi[4] = {0.f, 0.f, 0.f, 0.f}, vg[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix < pend) {
            int w = (int)(pix % W); long long t = pix / W; int h = (int)(t % H); int n = (int)(t / H);
            int co = cot * TC + lc;
            if (co < C) {
                uint2 rh = *reinterpret_cast<const uint2*>(gout + split_index(n, h, 0, w, co, H, W, C));
                uint2 rl = *reinterpret_cast<const uint2*>(gout + split_index(n, h, 1, w, co, H, W, C));
                const __nv_bfloat16* ph = reinterpret_cast<const __nv_bfloat16*>(&rh);
                const __nv_bfloat16* pl = reinterpret_cast<const __nv_bfloat16*>(&rl);
#pragma unroll
                for (int k = 0; k < 4; ++k) vg[k] = __bfloat162float(ph[k]) + __bfloat162float(pl[k]);
            }
            int ih = h + dh, iw = w + dw, ci = cit * TC + lc;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W && ci < C) {
                uint2 rh = *reinterpret_cast<const uint2*>(in + split_index(n, ih, 0, iw, ci, H, W, C));
                uint2 rl = *reinterpret_cast<const uint2*>(in + split_index(n, ih, 1, iw, ci, H, W, C));
                const __nv_bfloat16* ph = reinterpret_cast<const __nv_bfloat16*>(&rh);
                const __nv_bfloat16* pl = reinterpret_cast<const __nv_bfloat16*>(&rl);
#pragma unroll
                for (int k = 0; k < 4; ++k) vi[k] = __bfloat162float(ph[k]) + __bfloat162float(pl[k]);
            }
        }
        *reinterpret_cast<float4*>(&s_i[lp][lc]) = make_float4(vi[0], vi[1], vi[2], vi[3]);
        *reinterpret_cast<float4*>(&s_g[lp][lc]) = make_float4(vg[0], vg[1], vg[2], vg[3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&s_i[k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&s_g[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }
    float* dst = partial + (size_t)blockIdx.y * 9 * C * C + (size_t)tap * C * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int ci = cit * TC + ty * 4 + i;
        if (ci >= C) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = cot * TC + tx * 4 + j;
            if (co < C) dst[(size_t)ci * C + co] = acc[i][j];
        }
    }
}

int wgrad_simt_nparts(ConvShape s) {
    long long P = (long long)s.B * s.H * s.W;
    long long n = P / 1024;
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}

int launch_wgrad3x3_simt(const __nv_bfloat16* split_gout, const __nv_bfloat16* split_in, float* partial,
                         int* nparts_out, ConvShape s, cudaStream_t st) {
    if (s.C % 4 != 0) { set_error("SIMT wgrad: channels must be a multiple of 4 (got %d)", s.C); return -1; }
    int nparts = wgrad_simt_nparts(s);
    long long P = (long long)s.B * s.H * s.W;
    long long per = (P + nparts - 1) / nparts;
    per = (per + TK - 1) / TK * TK;
    int ctiles = (s.C + TC - 1) / TC;
    dim3 grid((unsigned)(9 * ctiles * ctiles), (unsigned)nparts);
    wgrad3x3_simt_kernel<<<grid, 256, 0, st>>>(split_gout, split_in, partial, s.B, s.H, s.W, s.C, per);
    count_launch();
    *nparts_out = nparts;
    return check_cuda(cudaGetLastError(), "wgrad3x3_simt launch");
}

}  // namespace msb
