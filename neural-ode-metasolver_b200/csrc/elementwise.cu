// Memory-bound helper kernels: activation+split prologue, weight packing, wgrad partial reduction.
// All are HBM/L2-bound streaming kernels: 128-bit accesses, grid = multiple of the SM count.
#include "msb_internal.h"

namespace msb {

// ---------------------------------------------------------------------------------------------
// act_split: split[B][H][2][W][C] = hi/lo(act(x) * mul * scale), dact = act'(x).   x, mul fp32 NHWC (mul optional).
// One thread handles 4 consecutive channels (float4 in, 2 x 8-byte bf16x4 out).
// ---------------------------------------------------------------------------------------------
struct SliceScales { float s[kMaxSlices]; };
__global__ void __launch_bounds__(256) act_split_kernel(const float* __restrict__ x, const float* __restrict__ mul, int act,
                                                        SliceScales scales, size_t vec_per_slice,
                                                        __nv_bfloat16* __restrict__ split,
                                                        float* __restrict__ dact, size_t n_vec, int W, int C) {
    const int cv = C >> 2;  // float4 per pixel
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        size_t pix = i / cv;
        int c4 = (int)(i - pix * cv);
        size_t row = pix / W;  // = n*H + h
        int w = (int)(pix - row * W);
        float4 v = reinterpret_cast<const float4*>(x)[i];
        const float scale = scales.s[vec_per_slice ? i / vec_per_slice : 0];
        float a[4], d[4];
        act_both(act, v.x, a[0], d[0]);
        act_both(act, v.y, a[1], d[1]);
        act_both(act, v.z, a[2], d[2]);
        act_both(act, v.w, a[3], d[3]);
        if (mul) {
            float4 m = reinterpret_cast<const float4*>(mul)[i];
            a[0] = __fmul_rn(a[0], m.x); a[1] = __fmul_rn(a[1], m.y); a[2] = __fmul_rn(a[2], m.z); a[3] = __fmul_rn(a[3], m.w);
        }
        __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16(__fmul_rn(a[k], scale), hi[k], lo[k]);
        size_t o_hi = ((row * 2 + 0) * W + w) * (size_t)C + (size_t)c4 * 4;
        size_t o_lo = ((row * 2 + 1) * W + w) * (size_t)C + (size_t)c4 * 4;
        *reinterpret_cast<uint2*>(split + o_hi) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(split + o_lo) = *reinterpret_cast<const uint2*>(lo);
        if (dact) reinterpret_cast<float4*>(dact)[i] = make_float4(d[0], d[1], d[2], d[3]);
    }
}

void launch_act_split(const float* x, const float* mul, int act, float scale, __nv_bfloat16* split, float* dact,
                      int B, int H, int W, int C, cudaStream_t st) {
    launch_act_split_sliced(x, mul, act, &scale, 1, split, dact, B, H, W, C, st);
}

// per-slice scale: the batch is n_slices equal slices (stacked solver axis), slice s is scaled by scales[s]
void launch_act_split_sliced(const float* x, const float* mul, int act, const float* scales_host, int n_slices,
                             __nv_bfloat16* split, float* dact, int B, int H, int W, int C, cudaStream_t st) {
    SliceScales scales;
    for (int i = 0; i < kMaxSlices; ++i) scales.s[i] = scales_host[i < n_slices ? i : 0];
    size_t n_vec = (size_t)B * H * W * C / 4;
    size_t vec_per_slice = n_slices > 1 ? n_vec / n_slices : 0;
    int blocks = (int)std::min<size_t>((n_vec + 255) / 256, (size_t)num_sms() * 8);
    if (blocks < 1) blocks = 1;
    act_split_kernel<<<blocks, 256, 0, st>>>(x, mul, act, scales, vec_per_slice, split, dact, n_vec, W, C);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// SIMT-engine weight pack: OIHW fp32 -> [tap][c_in][c_out] fp32 (c_out contiguous).
// transpose: the input-gradient convolution = conv with W'[i][o][r][s] = W[o][i][2-r][2-s].
// `skip_in` drops leading input channels (MNIST ConcatConv2d: channel 0 is the time channel).
// ---------------------------------------------------------------------------------------------
__global__ void pack_w_simt_kernel(const float* __restrict__ w, float* __restrict__ out, int C, int Cin_total,
                                   int skip_in, int transpose) {
    int total = 9 * C * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int co = i % C;           // output channel of the packed conv
        int ci = (i / C) % C;     // input channel of the packed conv
        int tap = i / (C * C);
        int r = tap / 3, s = tap % 3;
        float v;
        if (!transpose) v = w[(((size_t)co * Cin_total + (ci + skip_in)) * 3 + r) * 3 + s];
        else v = w[(((size_t)ci * Cin_total + (co + skip_in)) * 3 + (2 - r)) * 3 + (2 - s)];
        out[i] = v;
    }
}

void launch_pack_w_simt(const float* w, float* out, int C, int Cin_total, int skip_in, int transpose, cudaStream_t st) {
    int total = 9 * C * C;
    pack_w_simt_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, out, C, Cin_total, skip_in, transpose);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// tcgen05-engine weight pack: OIHW fp32 -> bf16 "A tiles" of 128 rows x 64 k-elements (K-major,
// 128-byte rows, not swizzled in global memory: TMA applies the 128B swizzle on the way to smem).
//   C == 64 : tile index = tap                       rows 0..63 = hi(W[co=row]), rows 64..127 = lo(W[co=row-64])
//             but rows are permuted so that the hi and lo row of one output channel sit 16 lanes
//             apart inside the same 32-lane TMEM quadrant (the epilogue adds them with one shuffle):
//             row m: q = m/32, part = (m%32)/16, co = 16*q + m%16.
//   C == 128: tile index = ((tap*2 + chunk)*2 + part) rows = co 0..127, part 0 = hi, 1 = lo,
//             chunk = 64-wide slice of c_in.
// Element (row, k) of a tile = W[co][chunk*64 + k][r][s] (or the transposed/rotated weight).
// ---------------------------------------------------------------------------------------------
__global__ void pack_w_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int C, int transpose,
                                 int cin_total, int skip_in) {
    int chunks = C / 64;
    int tiles = (C == 64) ? 9 : 9 * chunks * 2;
    int total = tiles * 128 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int k = i & 63;
        int row = (i >> 6) & 127;
        int tile = i >> 13;
        int tap, chunk, part, co;
        if (C == 64) {
            tap = tile; chunk = 0;
            int q = row >> 5;
            part = (row & 31) >> 4;
            co = 16 * q + (row & 15);
        } else {
            part = tile & 1;
            chunk = (tile >> 1) % chunks;
            tap = tile / (2 * chunks);
            co = row;
        }
        int ci = chunk * 64 + k;
        int r = tap / 3, s = tap % 3;
        float v;
        // cin_total / skip_in: the weight tensor has cin_total input channels of which the first skip_in are not part of
        // this GEMM (ConcatConv2d: channel 0 is the time channel)
        if (!transpose) v = w[(((size_t)co * cin_total + ci + skip_in) * 3 + r) * 3 + s];
        else v = w[(((size_t)ci * cin_total + co + skip_in) * 3 + (2 - r)) * 3 + (2 - s)];
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        out[i] = part ? lo : hi;
    }
}

void launch_pack_w_tc_ex(const float* w, __nv_bfloat16* out, int C, int transpose, int cin_total, int skip_in, cudaStream_t st) {
    int tiles = (C == 64) ? 9 : 9 * (C / 64) * 2;
    int total = tiles * 128 * 64;
    pack_w_tc_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, out, C, transpose, cin_total, skip_in);
    count_launch();
}
void launch_pack_w_tc(const float* w, __nv_bfloat16* out, int C, int transpose, cudaStream_t st) {
    launch_pack_w_tc_ex(w, out, C, transpose, C, 0, st);
}

// ---------------------------------------------------------------------------------------------
// wgrad finalisation: grad_w[O][I][3][3] (+)= sum_{p < nparts} partial[p][tap][ci][co]
// Deterministic (fixed summation order), one thread per weight.
// ---------------------------------------------------------------------------------------------
// grad_w may have more input channels than the GEMM (`cin_total` > C: ConcatConv2d, whose first `skip_in`
// input channels are handled elsewhere): element (co, ci, r, s) goes to grad_w[co][ci + skip_in][r][s].
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ grad_w,
                                    int C, int accumulate, int cin_total, int skip_in) {
    int total = 9 * C * C;
    // thread = one element of the partial layout [tap][ci][co] (co fastest): the nparts reads of a warp are coalesced;
    // the sum runs over the parts in ascending order (fixed order: bitwise reproducible), four loads in flight
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
        const int co = j % C, ci = (j / C) % C, tap = j / (C * C);
        const int r = tap / 3, s = tap - 3 * r;
        float acc = 0.f;
        int p = 0;
        for (; p + 4 <= nparts; p += 4) {
            const float a0 = partial[(size_t)p * total + j], a1 = partial[(size_t)(p + 1) * total + j];
            const float a2 = partial[(size_t)(p + 2) * total + j], a3 = partial[(size_t)(p + 3) * total + j];
            acc += a0; acc += a1; acc += a2; acc += a3;
        }
        for (; p < nparts; ++p) acc += partial[(size_t)p * total + j];
        float* q = grad_w + (((size_t)co * cin_total + ci + skip_in) * 3 + r) * 3 + s;
        *q = accumulate ? *q + acc : acc;
    }
}

void launch_wgrad_reduce(const float* partial, int nparts, float* grad_w, int C, int accumulate, cudaStream_t st,
                         int cin_total, int skip_in) {
    int total = 9 * C * C;
    wgrad_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>(partial, nparts, grad_w, C, accumulate,
                                                             cin_total > 0 ? cin_total : C, skip_in);
    count_launch();
}

// out[i] (+)= in[i]   (tiny helper: accumulate partial buffers across wgrad launches)
__global__ void axpy_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n, int accumulate) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = accumulate ? dst[i] + src[i] : src[i];
}

}  // namespace msb
