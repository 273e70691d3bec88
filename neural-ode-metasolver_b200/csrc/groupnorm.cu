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
