// The callers either side of the ODE-block path, each as ONE streaming kernel:
//   * the elementwise steps of the adversarial attacks that drive the input-gradient path
//     (MegaAdversarial/src/attacks/fgsm.py:27-40, 93-102, pgd.py:28-53): signed step, eps-ball clamp, [0,1]
//     projection, (un)normalisation -- same fp32 operations in the same order as the reference's torch calls
//     (no contraction: every op is an explicit round-to-nearest intrinsic), so results are bit-identical;
//   * the SGD-momentum / weight-decay update of examples/cifar10/train_and_attack.py:98-99, 322 on ONE flat
//     parameter buffer, with the 1/world gradient average of the data-parallel all-reduce folded in.
//   * the training-input transform of sopa/src/models/odenet_cifar10/data.py:40-46 (random crop with zero padding, random
//     horizontal flip, ToTensor, Normalize) from a uint8 dataset resident in HBM into the stem's fp32 NHWC input.
// HBM-bound, tiny (3 x 32 x 32 images, 675 k parameters): the point is one launch instead of ~10 per step.
#include "metasolver_b200.h"
#include "msb_internal.h"

namespace msb {
namespace {

struct ChanConst { float v[4][MSB_ATTACK_MAX_CHANNELS]; };   // up to four per-channel constants

__device__ __forceinline__ float sign_step(float g, float step) {      // step * sign(g), sign(0) = 0, sign(nan) = nan*
    return g > 0.f ? step : (g < 0.f ? -step : __fmul_rn(step, g));    // (g == 0 -> +-0 * step; nan propagates)
}
__device__ __forceinline__ float clamp_minmax(float x, float lo, float hi) { return fmaxf(fminf(x, hi), lo); }   // torch.max(torch.min(x, hi), lo)
__device__ __forceinline__ float project01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }                        // torch.clamp(x, 0, 1)

// channel of flat element i: NCHW-contiguous (c = (i / HW) % C) or channels-last memory (c = i % C)
__device__ __forceinline__ int chan_of(size_t i, int C, int HW, int channels_last) {
    return channels_last ? (int)(i % C) : (int)((i / HW) % C);
}

// kind: see MSB_ATTACK_* in the header.  k.v[0] = mean / lower, [1] = std / upper, [2] = eps, [3] = alpha (per channel)
__global__ void __launch_bounds__(256) attack_step_kernel(int kind, const float* __restrict__ a, const float* __restrict__ g,
                                                          const float* __restrict__ ref, float* __restrict__ out, size_t n,
                                                          int C, int HW, int channels_last, float eps, float step,
                                                          int normalize_out, ChanConst k) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = chan_of(i, C, HW, channels_last);
        float r;
        switch (kind) {
            case MSB_ATTACK_UNNORMALIZE:        // (x - inv_mean) / inv_std      fgsm.py:27, pgd.py:28
            case MSB_ATTACK_NORMALIZE:          // (x - mean) / std
                r = __fdiv_rn(__fsub_rn(a[i], k.v[0][c]), k.v[1][c]);
                break;
            case MSB_ATTACK_FGSM_STEP: {        // normalize(project(x + eps * sign(g)))      fgsm.py:38-40
                r = project01(__fadd_rn(a[i], sign_step(g[i], eps)));
                if (normalize_out) r = __fdiv_rn(__fsub_rn(r, k.v[0][c]), k.v[1][c]);
                break;
            }
            case MSB_ATTACK_PGD_STEP: {         // project(clamp(x + lr * sign(g), x0 - eps, x0 + eps))    pgd.py:47-51
                const float x0 = ref[i];
                r = __fadd_rn(a[i], sign_step(g[i], step));
                r = clamp_minmax(r, __fsub_rn(x0, eps), __fadd_rn(x0, eps));
                r = project01(r);
                if (normalize_out) r = __fdiv_rn(__fsub_rn(r, k.v[0][c]), k.v[1][c]);
                break;
            }
            case MSB_ATTACK_FGSMR_INIT: {       // delta = clamp(eps - (2 eps) u, lower - x, upper - x)     fgsm.py:93-95
                const float x = ref[i], e = k.v[2][c];
                r = __fsub_rn(e, __fmul_rn(__fmul_rn(2.f, e), a[i]));
                r = clamp_minmax(r, __fsub_rn(k.v[0][c], x), __fsub_rn(k.v[1][c], x));
                break;
            }
            default: {                          // MSB_ATTACK_FGSMR_STEP: x + clamp(clamp(delta + alpha sign(g), -eps, eps), lower - x, upper - x)
                const float x = ref[i], e = k.v[2][c];                                                  // fgsm.py:100-105
                r = __fadd_rn(a[i], sign_step(g[i], k.v[3][c]));
                r = clamp_minmax(r, -e, e);
                r = clamp_minmax(r, __fsub_rn(k.v[0][c], x), __fsub_rn(k.v[1][c], x));
                if (normalize_out) r = __fadd_rn(x, r);          // flag reused: emit x + delta instead of delta
                break;
            }
        }
        out[i] = r;
    }
}

// torch.optim.SGD (momentum, weight decay, no dampening / nesterov), single tensor:
//   g = grad * grad_scale (+ wd * p);  buf = first ? g : momentum * buf + g;  p -= lr * buf
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ p, const float* __restrict__ grad,
                                                       float* __restrict__ buf, size_t n, float lr, float momentum,
                                                       float wd, float grad_scale, int first) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float w = p[i];
        float g = grad[i];
        if (grad_scale != 1.f) g = __fmul_rn(g, grad_scale);
        if (wd != 0.f) g = __fadd_rn(g, __fmul_rn(wd, w));
        float b = g;
        if (momentum != 0.f) {
            b = first ? g : __fadd_rn(__fmul_rn(momentum, buf[i]), g);
            buf[i] = b;
        }
        p[i] = __fsub_rn(w, __fmul_rn(lr, b));
    }
}

// <a, b> over n floats in two deterministic stages (fixed grid, fixed order; double accumulation):
// the reductions behind the gradients w.r.t. the Butcher coefficients (odeblock.cu).
constexpr int kDotBlocks = 592;      // 4 per SM
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n4,
                                                          double* __restrict__ partial) {
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
        acc += (double)x.x * y.x + (double)x.y * y.y + (double)x.z * y.z + (double)x.w * y.w;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// <a, b> with b of `period` elements repeated along a (a per-image map against a batch): n4, period4 in float4 units
__global__ void __launch_bounds__(256) dot_bcast_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                size_t n4, size_t period4, double* __restrict__ partial) {
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i % period4];
        acc += (double)x.x * y.x + (double)x.y * y.y + (double)x.z * y.z + (double)x.w * y.w;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) dot_final_kernel(const double* __restrict__ partial, int nblocks, double scale,
                                                        double* __restrict__ out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out += scale * sh[0];
}

int grid_for(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

size_t dot_scratch_bytes() { return kDotBlocks * sizeof(double); }
// *out += scale * <a, b>   (n a multiple of 4; scratch = dot_scratch_bytes())
void launch_dot_accumulate(const float* a, const float* b, size_t n, double scale, double* out, double* scratch, cudaStream_t st) {
    dot_partial_kernel<<<kDotBlocks, 256, 0, st>>>(a, b, n / 4, scratch);
    dot_final_kernel<<<1, 256, 0, st>>>(scratch, kDotBlocks, scale, out);
    count_launch(2);
}

// *out += scale * <a, tile(b)>   (b has `period` elements, n a multiple of period, both multiples of 4)
void launch_dot_bcast_accumulate(const float* a, const float* b, size_t n, size_t period, double scale, double* out,
                                 double* scratch, cudaStream_t st) {
    dot_bcast_partial_kernel<<<kDotBlocks, 256, 0, st>>>(a, b, n / 4, period / 4, scratch);
    dot_final_kernel<<<1, 256, 0, st>>>(scratch, kDotBlocks, scale, out);
    count_launch(2);
}

// Training-input pipeline of data.py:40-46 for a batch whose uint8 source images are resident in HBM:
// RandomCrop(size, padding) (zero padding of the uint8 image), RandomHorizontalFlip, ToTensor (/255), Normalize, written
// straight into the fp32 NHWC layout the stem kernel reads.  One thread per output pixel (3 channels, 12 B out, 3 B in).
namespace {
struct AugConst { float mean[MSB_ATTACK_MAX_CHANNELS], std[MSB_ATTACK_MAX_CHANNELS]; };
__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ img, const int64_t* __restrict__ index,
                                                      const int32_t* __restrict__ dx, const int32_t* __restrict__ dy,
                                                      const uint8_t* __restrict__ flip, int B, int H, int W, int C, int pad,
                                                      AugConst k, float* __restrict__ out) {
    const size_t npix = (size_t)B * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((size_t)W * H));
        const int64_t src = index ? index[b] : b;
        const int xc = (flip && flip[b]) ? W - 1 - x : x;                  // flip acts on the cropped image
        const int sx = xc + (dx ? dx[b] : pad) - pad, sy = y + (dy ? dy[b] : pad) - pad;   // crop window offset in the padded image
        const bool inside = sx >= 0 && sx < W && sy >= 0 && sy < H;
        const uint8_t* s = img + (((size_t)src * H + (inside ? sy : 0)) * W + (inside ? sx : 0)) * C;
        float* o = out + i * C;
        for (int c = 0; c < C; ++c) {
            const float v = inside ? __fdiv_rn((float)s[c], 255.f) : 0.f;                  // ToTensor: uint8 -> float, div(255)
            o[c] = __fdiv_rn(__fsub_rn(v, k.mean[c]), k.std[c]);                           // Normalize: sub_(mean).div_(std)
        }
    }
}
}  // namespace

}  // namespace msb

using namespace msb;

extern "C" {

int msb_augment_batch(const uint8_t* images_u8, const int64_t* index, const int32_t* dx, const int32_t* dy, const uint8_t* flip,
                      int batch, int height, int width, int channels, int padding, const float* mean, const float* std,
                      float* out_nhwc, void* cuda_stream) {
    if (batch < 0 || height < 1 || width < 1 || channels < 1 || channels > MSB_ATTACK_MAX_CHANNELS || padding < 0) {
        set_error("msb_augment_batch: bad geometry (batch=%d, %dx%dx%d, padding=%d)", batch, height, width, channels, padding);
        return -1;
    }
    if (batch == 0) return 0;
    if (!images_u8 || !out_nhwc || !mean || !std) { set_error("msb_augment_batch: null buffer"); return -1; }
    AugConst k;
    for (int c = 0; c < MSB_ATTACK_MAX_CHANNELS; ++c) { k.mean[c] = c < channels ? mean[c] : 0.f; k.std[c] = c < channels ? std[c] : 1.f; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    augment_kernel<<<grid_for((size_t)batch * height * width), 256, 0, st>>>(images_u8, index, dx, dy, flip, batch, height, width,
                                                                             channels, padding, k, out_nhwc);
    count_launch();
    return check_cuda(cudaGetLastError(), "augment_batch launch");
}

int msb_attack_step(int kind, const float* a, const float* grad, const float* ref, float* out, int64_t n_elements,
                    int channels, int hw, int channels_last, float eps, float step, int normalize_out,
                    const float* chan_consts, void* cuda_stream) {
    if (kind < MSB_ATTACK_UNNORMALIZE || kind > MSB_ATTACK_FGSMR_STEP) { set_error("msb_attack_step: unknown kind %d", kind); return -1; }
    if (n_elements < 0) { set_error("msb_attack_step: negative size"); return -1; }
    if (n_elements > 0 && (!a || !out)) { set_error("msb_attack_step: null buffer"); return -1; }
    if (channels < 1 || channels > MSB_ATTACK_MAX_CHANNELS || hw < 1 || n_elements % ((int64_t)channels * hw) != 0) {
        set_error("msb_attack_step: bad geometry (channels=%d, hw=%d, n=%lld)", channels, hw, (long long)n_elements);
        return -1;
    }
    const bool needs_grad = kind == MSB_ATTACK_FGSM_STEP || kind == MSB_ATTACK_PGD_STEP || kind == MSB_ATTACK_FGSMR_STEP;
    const bool needs_ref = kind == MSB_ATTACK_PGD_STEP || kind == MSB_ATTACK_FGSMR_INIT || kind == MSB_ATTACK_FGSMR_STEP;
    if (n_elements > 0 && ((needs_grad && !grad) || (needs_ref && !ref))) { set_error("msb_attack_step: kind %d needs grad / ref", kind); return -1; }
    const bool needs_consts = kind != MSB_ATTACK_FGSM_STEP && kind != MSB_ATTACK_PGD_STEP;
    if ((needs_consts || normalize_out) && !chan_consts) { set_error("msb_attack_step: per-channel constants missing"); return -1; }
    ChanConst k;
    for (int j = 0; j < 4; ++j)
        for (int c = 0; c < MSB_ATTACK_MAX_CHANNELS; ++c)
            k.v[j][c] = (chan_consts && c < channels) ? chan_consts[j * channels + c] : (j == 1 ? 1.f : 0.f);
    if (n_elements == 0) return 0;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    attack_step_kernel<<<grid_for((size_t)n_elements), 256, 0, st>>>(kind, a, grad, ref, out, (size_t)n_elements, channels, hw,
                                                                     channels_last, eps, step, normalize_out, k);
    count_launch();
    return check_cuda(cudaGetLastError(), "attack_step launch");
}

int msb_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                 float weight_decay, float grad_scale, int first_step, void* cuda_stream) {
    if (!params || !grads || n < 0 || (momentum != 0.f && !momentum_buf)) { set_error("msb_sgd_step: null buffer / negative size"); return -1; }
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    sgd_step_kernel<<<grid_for((size_t)n), 256, 0, st>>>(params, grads, momentum_buf, (size_t)n, lr, momentum, weight_decay,
                                                         grad_scale, first_step);
    count_launch();
    return check_cuda(cudaGetLastError(), "sgd_step launch");
}

}  // extern "C"
