// Network head of MetaNODE (sopa/src/models/odenet_cifar10/layers.py:390-392, 425): AdaptiveAvgPool2d((1,1)) + Flatten +
// Linear, and the cross-entropy loss of the training loop (examples/cifar10/train_and_attack.py:303-311), forward and
// backward -- the last ATen / cuBLAS kernels of a premetanode10 step.  Tiny HBM-bound kernels:
//   pool_fc_fwd : one CTA per image; thread c sums its channel over the H*W pixels of the NHWC map (coalesced 128-byte
//                 ... 512-byte rows), pooled = sum / HW, then warp w computes logits w, w + nwarps, ...
//   pool_fc_bwd : dpooled = dlogits W, dx = dpooled / HW broadcast over the pixels (one CTA per image);
//                 dW = dlogits^T pooled and db = sum_n dlogits in a second, fixed-order (deterministic) kernel
//   ce_fwd/bwd  : loss = mean_n (logsumexp(z_n) - z_n[y_n]),  dz = (softmax(z) - onehot(y)) * gout / B
#include "metasolver_b200.h"
#include "msb_internal.h"

namespace msb {
namespace {

constexpr int kMaxClasses = 32;

__global__ void pool_fc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ pooled, float* __restrict__ logits, int HW, int C, int K) {
    extern __shared__ float sp[];                       // C pooled values
    const int n = blockIdx.x;
    const float* xn = x + (size_t)n * HW * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains (latency), summed in a fixed order
        int p = 0;
        for (; p + 3 < HW; p += 4) {
            a0 += xn[(size_t)p * C + c]; a1 += xn[(size_t)(p + 1) * C + c];
            a2 += xn[(size_t)(p + 2) * C + c]; a3 += xn[(size_t)(p + 3) * C + c];
        }
        for (; p < HW; ++p) a0 += xn[(size_t)p * C + c];
        const float m = ((a0 + a1) + (a2 + a3)) / (float)HW;
        sp[c] = m;
        pooled[(size_t)n * C + c] = m;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = warp; k < K; k += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(sp[c], w[(size_t)k * C + c], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) logits[(size_t)n * K + k] = a + (b ? b[k] : 0.f);
    }
}

__global__ void pool_fc_bwd_x_kernel(const float* __restrict__ dlogits, const float* __restrict__ w, float* __restrict__ dx,
                                     int HW, int C, int K) {
    extern __shared__ float sd[];                       // C values: dpooled / HW
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < K; ++k) a = fmaf(dlogits[(size_t)n * K + k], w[(size_t)k * C + c], a);
        sd[c] = a / (float)HW;
    }
    __syncthreads();
    float4* out = reinterpret_cast<float4*>(dx + (size_t)n * HW * C);
    const int c4 = C >> 2, total = HW * c4;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = (i % c4) * 4;
        out[i] = make_float4(sd[c], sd[c + 1], sd[c + 2], sd[c + 3]);
    }
}

// dW[k][c] = sum_n dlogits[n][k] * pooled[n][c], db[k] = sum_n dlogits[n][k]: one WARP per output (lane l takes images
// l, l+32, ... in order, then a fixed xor-shuffle tree: bitwise reproducible, and 16 dependent FMAs per lane at B = 512
// instead of 512 per thread)
__global__ void __launch_bounds__(128) pool_fc_bwd_w_kernel(const float* __restrict__ dlogits, const float* __restrict__ pooled,
                                                            float* __restrict__ dw, float* __restrict__ db, int B, int C, int K) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= K * C + (db ? K : 0)) return;
    float a = 0.f;
    if (i < K * C) {
        const int k = i / C, c = i - k * C;
        for (int n = lane; n < B; n += 32) a = fmaf(dlogits[(size_t)n * K + k], pooled[(size_t)n * C + c], a);
    } else {
        const int k = i - K * C;
        for (int n = lane; n < B; n += 32) a += dlogits[(size_t)n * K + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
        if (i < K * C) dw[i] = a;
        else db[i - K * C] = a;
    }
}

// one CTA: thread n handles sample n (strided), block reduction in a fixed order
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ z, const int64_t* __restrict__ y, float* __restrict__ loss,
                                                     float* __restrict__ lse_out, int B, int K) {
    __shared__ float sh[256];
    float acc = 0.f;
    for (int n = threadIdx.x; n < B; n += 256) {
        const float* zn = z + (size_t)n * K;
        float m = zn[0];
        for (int k = 1; k < K; ++k) m = fmaxf(m, zn[k]);
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += expf(zn[k] - m);
        const float lse = m + logf(s);
        lse_out[n] = lse;
        acc += lse - zn[y[n]];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = sh[0] / (float)B;
}

__global__ void ce_bwd_kernel(const float* __restrict__ z, const int64_t* __restrict__ y, const float* __restrict__ lse,
                              const float* __restrict__ gout, float* __restrict__ dz, int B, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int n = i / K, k = i - n * K;
    const float p = expf(z[i] - lse[n]);
    dz[i] = (p - (y[n] == k ? 1.f : 0.f)) * (gout[0] / (float)B);
}

}  // namespace
}  // namespace msb

using namespace msb;

extern "C" {

int msb_pool_fc_forward(const float* x_nhwc, const float* w, const float* bias, float* pooled, float* logits, int batch,
                        int hw, int channels, int classes, void* cuda_stream) {
    if (!x_nhwc || !w || !pooled || !logits || batch < 1 || hw < 1 || channels < 1 || classes < 1) {
        set_error("msb_pool_fc_forward: bad arguments"); return -1;
    }
    const int threads = channels >= 256 ? 256 : (channels >= 128 ? 128 : 64);
    pool_fc_fwd_kernel<<<batch, threads, channels * sizeof(float), (cudaStream_t)cuda_stream>>>(x_nhwc, w, bias, pooled, logits, hw,
                                                                                                channels, classes);
    count_launch();
    return check_cuda(cudaGetLastError(), "pool_fc forward launch");
}

int msb_pool_fc_backward(const float* dlogits, const float* w, const float* pooled, float* dx_nhwc, float* dw, float* dbias,
                         int batch, int hw, int channels, int classes, void* cuda_stream) {
    if (!dlogits || !w || !pooled || batch < 1 || hw < 1 || channels < 4 || channels % 4 || classes < 1) {
        set_error("msb_pool_fc_backward: bad arguments (channels must be a multiple of 4)"); return -1;
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dx_nhwc) {
        pool_fc_bwd_x_kernel<<<batch, 256, channels * sizeof(float), st>>>(dlogits, w, dx_nhwc, hw, channels, classes);
        count_launch();
    }
    if (dw) {
        const int total = classes * channels + classes;
        pool_fc_bwd_w_kernel<<<(total * 32 + 127) / 128, 128, 0, st>>>(dlogits, pooled, dw, dbias, batch, channels, classes);
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "pool_fc backward launch");
}

int msb_cross_entropy_forward(const float* logits, const int64_t* labels, float* loss, float* lse, int batch, int classes,
                              void* cuda_stream) {
    if (!logits || !labels || !loss || !lse || batch < 1 || classes < 1 || classes > kMaxClasses * 1024) {
        set_error("msb_cross_entropy_forward: bad arguments"); return -1;
    }
    ce_fwd_kernel<<<1, 256, 0, (cudaStream_t)cuda_stream>>>(logits, labels, loss, lse, batch, classes);
    count_launch();
    return check_cuda(cudaGetLastError(), "cross_entropy forward launch");
}

int msb_cross_entropy_backward(const float* logits, const int64_t* labels, const float* lse, const float* grad_loss,
                               float* dlogits, int batch, int classes, void* cuda_stream) {
    if (!logits || !labels || !lse || !grad_loss || !dlogits || batch < 1 || classes < 1) {
        set_error("msb_cross_entropy_backward: bad arguments"); return -1;
    }
    const int total = batch * classes;
    ce_bwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)cuda_stream>>>(logits, labels, lse, grad_loss, dlogits, batch, classes);
    count_launch();
    return check_cuda(cudaGetLastError(), "cross_entropy backward launch");
}

}  // extern "C"
