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

// one warp per (sample, group); state is tiny (C=64, 6x6) so everything stays in L1/L2
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

int launch_groupnorm_epi(const float* x, const float* gamma, const float* beta, const EpiParams& epi, ConvShape s,
                         int groups, float eps, cudaStream_t st) {
    if (groups < 1 || s.C % groups) { set_error("groupnorm: %d channels not divisible into %d groups", s.C, groups); return -1; }
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
