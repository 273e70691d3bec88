// GroupNorm fused with the generic epilogue, and the ConcatConv2d "time channel" folding -- the extra
// pieces of the MNIST right-hand side (sopa/src/models/odenet_mnist/layers.py:158-171, 240-253).
//
//   y = (x - mean_g) * rstd_g * gamma[c] + beta[c]     per (sample, group), eps inside the sqrt
// followed by epilogue_apply(), so one launch yields either the bf16 hi/lo operand of the next
// convolution (norm -> ReLU -> split) or, for norm3, the Runge-Kutta stage combination.
// ConcatConv2d(t, x) = conv(cat[t*1, x]) = conv_{C->C}(x) + bias + t * tapmap, where
// tapmap[h][w][o] = sum over the in-bounds taps of W[o][0][r][s]   (the padding is zero, the t plane is not).
#include "msb_internal.h"

namespace msb {

// one warp per (sample, group); state is tiny (C=64, 6x6) so everything stays in L1/L2 (large states: the block kernels below)
__global__ void __launch_bounds__(128) groupnorm_epi_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, EpiParams epi, int B, int H,
                                                            int W, int C, int G, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B * G) return;
    const int n = warp / G, g = warp - n * G;
    const int cpg = C / G;
    const int HW = H * W;
    const int cnt = cpg * HW;
    const float* xb = x + (size_t)n * HW * C + (size_t)g * cpg;
    float s = 0.f;
    for (int i = lane; i < cnt; i += 32) s += xb[(size_t)(i / cpg) * C + (i % cpg)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)cnt;
    float v = 0.f;
    for (int i = lane; i < cnt; i += 32) {
        float d = xb[(size_t)(i / cpg) * C + (i % cpg)] - mean;
        v += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)cnt + eps);
    for (int i = lane; i < cnt; i += 32) {
        const int p = i / cpg, c = g * cpg + (i % cpg);
        const int h = p / W, w = p - h * W;
        const size_t idx = ((size_t)n * HW + p) * C + c;
        const float y = (x[idx] - mean) * rstd * gamma[c] + beta[c];
        epilogue_apply(epi, y, idx, n, h, w, c, H, W, C);
    }
}


// ---------------------------------------------------------------------------------------------
// Large states (CIFAR 'GN' / 'LN' / 'IN' right-hand sides: 64 x 32 x 32, 128 x 16 x 16): one CTA of 256 threads per SAMPLE,
// all groups at once; every warp instruction reads whole 128-byte NHWC lines, the per-thread partials are combined
// through shared memory in a fixed order (deterministic).  Same two-pass mean / variance as the warp kernel.
// Measured (B = 256, RK2 8 steps, fwd + bwd of the C = 64 block, scripts/bench_gn_rhs.py): GN 27.6 -> 13.6 ms,
// LN 177.8 -> 13.3 ms, IN 45.2 -> 13.3 ms (normalisation-free block: 6.3 ms).
// ---------------------------------------------------------------------------------------------
constexpr int kGnBlock = 256;

// ---- a thread owns 8 consecutive channels of a pixel (256-bit accesses, the
//      tensor-core engines' packed epilogue epi_finish_v8: same operations, same bits as epilogue_apply) ----
__device__ __forceinline__ void ld8(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// per-channel totals over the pixel lanes, then per-group totals; every thread gets the totals of its 8 channels' groups
// (and, optionally, of its channels).  sh: PL * C + 2 * C floats.
__device__ __forceinline__ void gn_reduce8(const float* v, float* grp_out, float* chan_out, float* sh, int C, int G, int PL) {
    const int t = threadIdx.x, CL = C >> 3, cl = t % CL, pl = t / CL, cpg = C / G;
    float* sh_col = sh + PL * C;
    float* sh_grp = sh_col + C;
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[pl * C + cl * 8 + j] = v[j];
    __syncthreads();
    if (t < C) {
        float a = 0.f;
        for (int q = 0; q < PL; ++q) a += sh[q * C + t];
        sh_col[t] = a;
    }
    __syncthreads();
    if (t < G) {
        float a = 0.f;
        for (int cc = 0; cc < cpg; ++cc) a += sh_col[t * cpg + cc];
        sh_grp[t] = a;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        grp_out[j] = sh_grp[(cl * 8 + j) / cpg];
        if (chan_out) chan_out[j] = sh_col[cl * 8 + j];
    }
    __syncthreads();
}

template <int ACT>
__global__ void __launch_bounds__(kGnBlock) groupnorm_epi_block8_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, EpiParams epi, int H, int W,
                                                                        int C, int G, float eps) {
    extern __shared__ float sh[];
    const int n = blockIdx.x, t = threadIdx.x;
    const int CL = C >> 3, cl = t % CL, pl = t / CL, PL = kGnBlock / CL, HW = H * W;
    const float cnt = (float)((C / G) * HW);
    const float* xb = x + (size_t)n * HW * C + cl * 8;
    float s[8], mean[8], rstd[8], xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll 4
    for (int p = pl; p < HW; p += PL) {
        ld8(xb + (size_t)p * C, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += xv[j];
    }
    gn_reduce8(s, mean, nullptr, sh, C, G, PL);
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean[j] /= cnt; s[j] = 0.f; }
#pragma unroll 4
    for (int p = pl; p < HW; p += PL) {
        ld8(xb + (size_t)p * C, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = xv[j] - mean[j]; s[j] += d * d; }
    }
    gn_reduce8(s, rstd, nullptr, sh, C, G, PL);
    float gm[8], bt[8];
    ld8(gamma + cl * 8, gm);
    ld8(beta + cl * 8, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) rstd[j] = rsqrtf(rstd[j] / cnt + eps);
    const EpiCoef coef = epi_coef(epi, n);
    const size_t plane_stride = (size_t)W * C;
    EpiVec8 ops;
    for (int p = pl; p < HW; p += PL) {
        const int h = p / W, w = p - h * W;
        const size_t idx = ((size_t)n * HW + p) * C + cl * 8;
        epi_prefetch_vec8(epi, idx, ops);
        ld8(x + idx, xv);
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (xv[j] - mean[j]) * rstd[j] * gm[j] + bt[j];
        epi_finish_v8<ACT>(epi, coef, y, ops, idx, split_index(n, h, 0, w, cl * 8, H, W, C), plane_stride);
    }
}

template <int ACT>
__global__ void __launch_bounds__(kGnBlock) groupnorm_bwd_epi_block8_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                            const float* __restrict__ beta, const float* __restrict__ dy,
                                                                            float dy_scale, int relu, EpiParams epi,
                                                                            float* __restrict__ dgamma_part, float* __restrict__ dbeta_part,
                                                                            int accumulate, int H, int W, int C, int G, float eps) {
    extern __shared__ float sh[];
    const int n = blockIdx.x, t = threadIdx.x;
    const int CL = C >> 3, cl = t % CL, pl = t / CL, PL = kGnBlock / CL, HW = H * W;
    const float cnt = (float)((C / G) * HW);
    const float* xb = x + (size_t)n * HW * C + cl * 8;
    const float* dyb = dy + (size_t)n * HW * C + cl * 8;
    float s[8], mean[8], rstd[8], xv[8], gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll 4
    for (int p = pl; p < HW; p += PL) {
        ld8(xb + (size_t)p * C, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += xv[j];
    }
    gn_reduce8(s, mean, nullptr, sh, C, G, PL);
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean[j] /= cnt; s[j] = 0.f; }
#pragma unroll 4
    for (int p = pl; p < HW; p += PL) {
        ld8(xb + (size_t)p * C, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = xv[j] - mean[j]; s[j] += d * d; }
    }
    gn_reduce8(s, rstd, nullptr, sh, C, G, PL);
    float gm[8], bt[8];
    ld8(gamma + cl * 8, gm);
    ld8(beta + cl * 8, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) rstd[j] = rsqrtf(rstd[j] / cnt + eps);
    auto upstream = [&](const float* xv_, const float* dy_, float* xh, float* gg) {       // xhat and g = dy * scale * act'(y)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            xh[j] = (xv_[j] - mean[j]) * rstd[j];
            const float y = (xv_[j] - mean[j]) * rstd[j] * gm[j] + bt[j];
            float g = dy_[j] * dy_scale;
            if (relu == ACT_RELU) { if (!(y > 0.f)) g = 0.f; }
            else if (relu != ACT_NONE) g *= dact_f(relu, y);
            gg[j] = g;
        }
    };
    float a[8], b[8], xh[8], gg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
#pragma unroll 2
    for (int p = pl; p < HW; p += PL) {
        ld8(xb + (size_t)p * C, xv);
        ld8(dyb + (size_t)p * C, gv);
        upstream(xv, gv, xh, gg);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += gg[j]; b[j] += gg[j] * xh[j]; }
    }
    float m1[8], m2[8], ta[8], tb[8];
    // the per-channel totals of a and b are the parameter-gradient partials of this sample; then the group sums of
    // gamma * a and gamma * b
    {
        float ga[8], gb[8];
        gn_reduce8(a, m1, ta, sh, C, G, PL);          // (group sums of plain a / b are not needed: m1, m2 are overwritten below)
        gn_reduce8(b, m2, tb, sh, C, G, PL);
        if (pl == 0 && dgamma_part) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const size_t q = (size_t)n * C + cl * 8 + j;
                dbeta_part[q] = accumulate ? dbeta_part[q] + ta[j] : ta[j];
                dgamma_part[q] = accumulate ? dgamma_part[q] + tb[j] : tb[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { ga[j] = gm[j] * a[j]; gb[j] = gm[j] * b[j]; }
        gn_reduce8(ga, m1, nullptr, sh, C, G, PL);
        gn_reduce8(gb, m2, nullptr, sh, C, G, PL);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { m1[j] /= cnt; m2[j] /= cnt; }
    const EpiCoef coef = epi_coef(epi, n);
    const size_t plane_stride = (size_t)W * C;
    EpiVec8 ops;
    for (int p = pl; p < HW; p += PL) {
        const int h = p / W, w = p - h * W;
        const size_t idx = ((size_t)n * HW + p) * C + cl * 8;
        epi_prefetch_vec8(epi, idx, ops);
        ld8(x + idx, xv);
        ld8(dy + idx, gv);
        upstream(xv, gv, xh, gg);
        float dx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] = rstd[j] * (gg[j] * gm[j] - m1[j] - xh[j] * m2[j]);
        epi_finish_v8<ACT>(epi, coef, dx, ops, idx, split_index(n, h, 0, w, cl * 8, H, W, C), plane_stride);
    }
}

__host__ inline bool gn_block8_form(const EpiParams& e, int C, int HW, int G) {
    const int CL = C / 8;
    return C % 8 == 0 && CL <= kGnBlock && kGnBlock % CL == 0 && C <= kGnBlock && HW >= 256 && HW % (kGnBlock / CL) == 0 &&
           !e.chan_bias && !e.pix_bias;
}
inline size_t gn_block8_smem(int C) { return ((size_t)(kGnBlock / (C / 8)) * C + 2 * C) * sizeof(float); }

int launch_groupnorm_epi(const float* x, const float* gamma, const float* beta, const EpiParams& epi, ConvShape s,
                         int groups, float eps, cudaStream_t st) {
    if (groups < 1 || s.C % groups) { set_error("groupnorm: %d channels not divisible into %d groups", s.C, groups); return -1; }
    if (tune_get(TUNE_GN_BLOCK) == 1 && gn_block8_form(epi, s.C, s.H * s.W, groups)) {
        const size_t smem = gn_block8_smem(s.C);
        if (epi.act == ACT_GELU) groupnorm_epi_block8_kernel<ACT_GELU><<<s.B, kGnBlock, smem, st>>>(x, gamma, beta, epi, s.H, s.W, s.C, groups, eps);
        else if (epi.act == ACT_RELU) groupnorm_epi_block8_kernel<ACT_RELU><<<s.B, kGnBlock, smem, st>>>(x, gamma, beta, epi, s.H, s.W, s.C, groups, eps);
        else groupnorm_epi_block8_kernel<ACT_NONE><<<s.B, kGnBlock, smem, st>>>(x, gamma, beta, epi, s.H, s.W, s.C, groups, eps);
        count_launch();
        return check_cuda(cudaGetLastError(), "groupnorm launch");
    }
    const int warps = s.B * groups;
    groupnorm_epi_kernel<<<(warps * 32 + 127) / 128, 128, 0, st>>>(x, gamma, beta, epi, s.B, s.H, s.W, s.C, groups, eps);
    count_launch();
    return check_cuda(cudaGetLastError(), "groupnorm launch");
}

// ---------------------------------------------------------------------------------------------
// GroupNorm backward, fused with the derivative of the activation that followed it in the forward pass (`relu` is an
// activation code: 0 = none, ACT_GELU, ACT_RELU) and with the generic epilogue:
//   xhat = (x - mean) rstd,  y = xhat gamma + beta,  g = dy * dy_scale * act'(y)
//   dgamma_part[n][c] (+)= sum_p g xhat        dbeta_part[n][c] (+)= sum_p g          (per-sample partials:
//                                               every (n, c) is owned by one warp -> deterministic)
//   dx = rstd (g gamma - mean_grp(g gamma) - xhat mean_grp(g gamma xhat))  ->  epilogue_apply(dx)
// mean / rstd are recomputed exactly as in groupnorm_epi_kernel, so y and the mask equal the forward bits.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) groupnorm_bwd_epi_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ dy,
                                                                float dy_scale, int relu, EpiParams epi,
                                                                float* __restrict__ dgamma_part, float* __restrict__ dbeta_part,
                                                                int accumulate, int B, int H, int W, int C, int G, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B * G) return;
    const int n = warp / G, g = warp - n * G;
    const int cpg = C / G;
    const int HW = H * W;
    const int cnt = cpg * HW;
    const float* xb = x + (size_t)n * HW * C + (size_t)g * cpg;
    float s = 0.f;
    for (int i = lane; i < cnt; i += 32) s += xb[(size_t)(i / cpg) * C + (i % cpg)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)cnt;
    float v = 0.f;
    for (int i = lane; i < cnt; i += 32) {
        float d = xb[(size_t)(i / cpg) * C + (i % cpg)] - mean;
        v += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)cnt + eps);
    // per-channel sums, then the two group sums
    float s1 = 0.f, s2 = 0.f;
    for (int cc = 0; cc < cpg; ++cc) {
        const int c = g * cpg + cc;
        const float gm = gamma[c], bt = beta[c];
        float a = 0.f, b = 0.f;
        for (int p = lane; p < HW; p += 32) {
            const size_t idx = ((size_t)n * HW + p) * C + c;
            const float xh = (x[idx] - mean) * rstd;
            const float y = (x[idx] - mean) * rstd * gm + bt;
            float gg = dy[idx] * dy_scale;
            if (relu == ACT_RELU) { if (!(y > 0.f)) gg = 0.f; }
            else if (relu != ACT_NONE) gg *= dact_f(relu, y);
            a += gg;
            b += gg * xh;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0 && dgamma_part) {
            const size_t q = (size_t)n * C + c;
            dbeta_part[q] = accumulate ? dbeta_part[q] + a : a;
            dgamma_part[q] = accumulate ? dgamma_part[q] + b : b;
        }
        s1 += gm * a;
        s2 += gm * b;
    }
    const float m1 = s1 / (float)cnt, m2 = s2 / (float)cnt;
    for (int i = lane; i < cnt; i += 32) {
        const int p = i / cpg, c = g * cpg + (i % cpg);
        const int h = p / W, w = p - h * W;
        const size_t idx = ((size_t)n * HW + p) * C + c;
        const float gm = gamma[c];
        const float xh = (x[idx] - mean) * rstd;
        const float y = (x[idx] - mean) * rstd * gm + beta[c];
        float gg = dy[idx] * dy_scale;
        if (relu == ACT_RELU) { if (!(y > 0.f)) gg = 0.f; }
        else if (relu != ACT_NONE) gg *= dact_f(relu, y);
        const float dx = rstd * (gg * gm - m1 - xh * m2);
        epilogue_apply(epi, dx, idx, n, h, w, c, H, W, C);
    }
}


int launch_groupnorm_bwd_epi(const float* x, const float* gamma, const float* beta, const float* dy, float dy_scale, int relu,
                             const EpiParams& epi, float* dgamma_part, float* dbeta_part, int accumulate, ConvShape s,
                             int groups, float eps, cudaStream_t st) {
    if (groups < 1 || s.C % groups) { set_error("groupnorm: %d channels not divisible into %d groups", s.C, groups); return -1; }
    if (tune_get(TUNE_GN_BLOCK) == 1 && gn_block8_form(epi, s.C, s.H * s.W, groups)) {
        const size_t smem = gn_block8_smem(s.C);
#define MSB_GN_BWD8(A) groupnorm_bwd_epi_block8_kernel<A><<<s.B, kGnBlock, smem, st>>>(x, gamma, beta, dy, dy_scale, relu, epi, \
                           dgamma_part, dbeta_part, accumulate, s.H, s.W, s.C, groups, eps)
        if (epi.act == ACT_GELU) MSB_GN_BWD8(ACT_GELU);
        else if (epi.act == ACT_RELU) MSB_GN_BWD8(ACT_RELU);
        else MSB_GN_BWD8(ACT_NONE);
#undef MSB_GN_BWD8
        count_launch();
        return check_cuda(cudaGetLastError(), "groupnorm backward launch");
    }
    const int warps = s.B * groups;
    groupnorm_bwd_epi_kernel<<<(warps * 32 + 127) / 128, 128, 0, st>>>(x, gamma, beta, dy, dy_scale, relu, epi, dgamma_part,
                                                                        dbeta_part, accumulate, s.B, s.H, s.W, s.C, groups, eps);
    count_launch();
    return check_cuda(cudaGetLastError(), "groupnorm backward launch");
}

// out[c] = sum_n part[n][c]   (fixed order)
__global__ void sum_over_batch_kernel(const float* __restrict__ part, float* __restrict__ out, int B, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f;
    for (int n = 0; n < B; ++n) a += part[(size_t)n * C + c];
    out[c] = a;
}
void launch_sum_over_batch(const float* part, float* out, int B, int C, cudaStream_t st) {
    sum_over_batch_kernel<<<(C + 127) / 128, 128, 0, st>>>(part, out, B, C);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// ConcatConv2d parameter gradients that do not come from the x-channel weight-gradient GEMM:
//   gb[o] (+)= sum_{n,p} dP[n][p][o]
//   gw[o][0][r][s] (+)= t * sum_{n, (h,w) : (h+r-1, w+s-1) inside the image} dP[n][h][w][o]     (time channel)
// One block per output channel, fixed-order tree reduction (deterministic).  gw: [C][C+1][3][3].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) concat_aux_grad_kernel(const float* __restrict__ dP, float t, float* __restrict__ gw,
                                                              float* __restrict__ gb, int accumulate, int B, int H, int W, int C) {
    __shared__ float red[10][128];
    const int o = blockIdx.x;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.f;
    const int total = B * H * W;
    for (int i = threadIdx.x; i < total; i += 128) {
        const int w = i % W, h = (i / W) % H;
        const float v = dP[(size_t)i * C + o];
        acc[9] += v;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ih = h + r - 1, iw = w + s - 1;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) acc[r * 3 + s] += v;
            }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) red[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int step = 64; step > 0; step >>= 1) {
        if (threadIdx.x < step)
#pragma unroll
            for (int k = 0; k < 10; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + step];
        __syncthreads();
    }
    if (threadIdx.x < 9) {
        float* q = gw + (size_t)o * (C + 1) * 9 + threadIdx.x;
        const float v = t * red[threadIdx.x][0];
        *q = accumulate ? *q + v : v;
    } else if (threadIdx.x == 9) {
        gb[o] = accumulate ? gb[o] + red[9][0] : red[9][0];
    }
}
void launch_concat_aux_grad(const float* dP, float t, float* gw, float* gb, int accumulate, ConvShape s, cudaStream_t st) {
    concat_aux_grad_kernel<<<s.C, 128, 0, st>>>(dP, t, gw, gb, accumulate, s.B, s.H, s.W, s.C);
    count_launch();
}

// tapmap[h][w][o] from the time-channel weights W[o][0][3][3] of a [C][C+1][3][3] ConcatConv2d weight
__global__ void time_tapmap_kernel(const float* __restrict__ w, float* __restrict__ tapmap, int H, int W, int C) {
    const int total = H * W * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int o = i % C, p = i / C;
        const int hh = p / W, ww = p - hh * W;
        float acc = 0.f;
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                const int ih = hh + r - 1, iw = ww + s - 1;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) acc += w[(((size_t)o * (C + 1)) * 3 + r) * 3 + s];
            }
        tapmap[i] = acc;
    }
}

void launch_time_tapmap(const float* w, float* tapmap, int H, int W, int C, cudaStream_t st) {
    const int total = H * W * C;
    time_tapmap_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, tapmap, H, W, C);
    count_launch();
}

}  // namespace msb
