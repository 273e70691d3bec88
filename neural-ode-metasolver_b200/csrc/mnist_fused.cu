// MNIST ODE block (sopa/src/models/odenet_mnist/layers.py:158-171, 250-253), forward, as ONE persistent launch:
// the whole fixed-step Runge-Kutta solve of an image -- every stage of every step -- runs inside one CTA with the
// state, the stage derivatives and both convolution operands resident in registers / shared / tensor memory.
//
//   f(t, x) = GN3(cconv2(t, relu(GN2(cconv1(t, relu(GN1(x)))))))        cconv(t, a) = conv3x3(a) + bias + t * tapmap
//
// (time channel of ConcatConv2d folded into a per-pixel map, see groupnorm.cu).  The state of an image is 64 ch x 6 x 6
// = 9 KB; round 1 ran ~25 launches per evaluation x 16 evaluations on the fp32 SIMT engine (1.9 ms for B = 128).
//
// One CTA = one image at a time (persistent over images).  Both convolutions are tcgen05 implicit GEMMs in the
// channel-major orientation  D[(c_out, hi/lo) = 128 lanes][48 padded pixels]:
//   conv1: A = [W1_hi ; W1_lo] resident in TENSOR MEMORY (288 columns, written once), B from shared memory;
//   conv2: A = [W2_hi ; W2_lo] resident in SHARED memory (nine 16 KB K-major tiles), B from shared memory
//          (the second weight set does not fit beside the first in the 512 TMEM columns);
//   B     = the activation image, bf16 hi / lo planes, zero-padded to 8 x 8 pixels so that a vertical tap is an offset of
//           8 pixels = 1024 B (one swizzle atom) and a window of 48 consecutive pixels covers the 6 output rows (columns 6, 7
//           of a row are never used); the horizontal taps read three copies shifted by 0 / 1 / 2 pixels.  K = (tap, c_in);
//           the X_hi and X_lo planes are separate MMAs into the same accumulator -> all four hi/lo products.
// Epilogue threads (128 = 4 warps, warp = TMEM lane quadrant): lane l < 16 holds the W_hi row of channel 16 q + l, lane
// l + 16 its W_lo row; one shuffle adds them, then the pair splits the image: lane < 16 keeps image rows 0..2, lane >= 16
// rows 3..5 (18 pixels of one channel per thread).  A GroupNorm group (2 channels x 36 pixels) is four threads: the
// statistics are two shuffles.  State y, stage input x_i and the stage derivatives k_j live in registers.
// With a tape (training) the stage input X, the convolution outputs P1 / P2 and the operands A / Hs are also written to
// their slots for the (multi-launch, SIMT) backward pass; the arithmetic is otherwise that of the unfused path.
#include <cuda.h>

#include "metasolver_b200.h"
#include "msb_internal.h"
#include "msb_host.h"
#include "msb_ptx.cuh"

namespace msb {

void launch_pack_w_tct_ex(const float* w, void* out, int transpose, int cin_total, int skip_in, cudaStream_t st);
void launch_pack_w_tc_ex(const float* w, __nv_bfloat16* out, int C, int transpose, int cin_total, int skip_in, cudaStream_t st);
void launch_time_tapmap(const float* w, float* tapmap, int H, int W, int C, cudaStream_t st);

namespace {

constexpr int kC = 64, kH = 6, kW = 6, kHW = 36;
constexpr int kComputeWarps = 4;
constexpr int kThreads = (kComputeWarps + 1) * 32;           // + the MMA-issuing warp
constexpr int kMaxStepsFused = 64;
constexpr int kPx = 18;                                      // pixels per thread (half an image)
constexpr int kNWin = 48;                                    // GEMM N: 6 output rows x 8 padded columns
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kD1Col = 0, kD2Col = 64, kW1Col = 128;    // accumulators (48 of 64 columns each), W1 image (288 columns)
constexpr int kImgBytes = 64 * 128;                          // one padded 8 x 8 image plane: 64 pixels x 64 bf16
constexpr int kActBytes = 3 * 2 * kImgBytes;                 // 3 horizontal shifts x {hi, lo}
constexpr int kW2Bytes = 9 * 128 * 128;                      // nine K-major tiles of 128 rows x 64 bf16
constexpr int kMapFloats = kHW * kC;

struct FusedParams {
    const float* x; float* y_out;
    const uint4* w1_tmem_img;          // pack_w_tct layout (18 chunks x 128 lanes x 16 words)
    const __nv_bfloat16* w2_tiles;     // pack_w_tc layout (9 tiles x 128 rows x 64)
    const float* tapmap[2];            // [36][64]
    const float* conv_b[2];
    const float* gamma[3]; const float* beta[3];
    float eps;
    int B, S, N;
    float b[MSB_MAX_STAGES], w[MSB_MAX_STAGES * MSB_MAX_STAGES];
    float dt[kMaxStepsFused];
    float ts[kMaxStepsFused * MSB_MAX_STAGES];     // stage times t_n + c_i dt (host-rounded like the reference)
    // tape (nullptr = inference): slot s = step * S + stage, 5 arrays of `slot_stride` bytes each: X, P1, P2 (fp32), A, Hs (split)
    char* tape; size_t slot_stride;
};

struct __align__(8) FBars { uint64_t mma_done; uint32_t tmem_base; };

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// GroupNorm over a group = 2 adjacent channels x 36 pixels = the 4 threads (lane, lane^1, lane^16, lane^17); v[] holds this
// thread's 18 pixels of its channel.  Two-pass (mean, then centred sum of squares), fp32, as groupnorm_epi_kernel.
__device__ __forceinline__ void group_stats(const float* v, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kPx; ++j) s += v[j];
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    mean = s / 72.f;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kPx; ++j) { const float d = v[j] - mean; q += d * d; }
    q += __shfl_xor_sync(0xffffffffu, q, 16);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    rstd = rsqrtf(q / 72.f + eps);
}

template <int S>
__global__ void __launch_bounds__(kThreads, 1) mnist_fused_fwd_kernel(const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_w2 = smem;                                   // 147 456 B
    uint8_t* smem_act = smem + kW2Bytes;                       // 49 152 B: [shift][plane][64 px][128 B], 128B-swizzled
    float* smem_map = reinterpret_cast<float*>(smem_act + kActBytes);     // 2 x [36][64] time-channel tap maps
    FBars* bars = reinterpret_cast<FBars*>(smem_map + 2 * kMapFloats);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(&bars->mma_done, 1); ptx::fence_barrier_init(); }
    if (warp == kComputeWarps) { ptx::tmem_alloc(&bars->tmem_base, kTmemCols); ptx::tmem_relinquish(); }
    // one-time loads: W2 tiles (plain copies; they were packed K-major WITHOUT swizzle -> swizzle here), zeroed activation
    // image (the padding never changes), tap maps
    for (int i = threadIdx.x; i < kActBytes / 16; i += kThreads) reinterpret_cast<uint4*>(smem_act)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 2 * kMapFloats; i += kThreads)
        smem_map[i] = (i < kMapFloats ? p.tapmap[0][i] : p.tapmap[1][i - kMapFloats]);
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.w2_tiles);
        for (int i = threadIdx.x; i < kW2Bytes / 16; i += kThreads) {
            const int row = i >> 3, chunk = i & 7;             // row of 128 B within the 9 x 128 rows, 16-byte chunk
            reinterpret_cast<uint4*>(smem_w2)[row * 8 + (chunk ^ (row & 7))] = __ldg(src + i);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tb = bars->tmem_base;
    if (warp < kComputeWarps) {            // W1 -> TMEM: thread = lane of the quadrant, 18 chunks of 16 columns
        const int L = warp * 32 + lane;
        const uint32_t w_lane = tb + ((uint32_t)(warp * 32) << 16) + kW1Col;
        for (int ch = 0; ch < 18; ++ch) {
            const uint4* src = p.w1_tmem_img + ((size_t)ch * 128 + L) * 4;
            uint32_t r[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) { const uint4 t = __ldg(src + v); r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w; }
            tmem_st16(w_lane + (uint32_t)ch * 16u, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_async_smem();                    // generic-proxy writes of W2 / zeros -> visible to the tensor core's reads
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == kComputeWarps) {
        // ===================== MMA issuer: waits for "operand image ready" (barrier 1), issues one convolution, commits =====================
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const uint32_t act_u32 = ptx::smem_u32(smem_act), w2_u32 = ptx::smem_u32(smem_w2);
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kNWin, 0, 0);
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        for (int n = blockIdx.x; n < p.B; n += gridDim.x)
            for (int ev = 0; ev < p.N * S; ++ev)
                for (int conv = 0; conv < 2; ++conv) {
                    bar_sync(1, kThreads);                     // the compute warps have written the operand image
                    ptx::tc_fence_after();
                    if (leader) {
                        const uint32_t d = tbu + (conv ? kD2Col : kD1Col);
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int s = 0; s < 3; ++s)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
#pragma unroll
                                    for (int pl = 0; pl < 2; ++pl) {
                                        const uint64_t bdesc = ptx::make_smem_desc_sw128(
                                            act_u32 + (uint32_t)((s * 2 + pl) * kImgBytes + r * 1024 + k * 32), 16, 1024);
                                        const uint32_t acc = (r | s | k | pl) ? 1u : 0u;
                                        if (conv == 0) {
                                            umma_ts(d, tbu + kW1Col + (uint32_t)((r * 3 + s) * 32 + k * 8), bdesc, idesc, acc);
                                        } else {
                                            const uint64_t adesc = ptx::make_smem_desc_sw128(w2_u32 + (uint32_t)((r * 3 + s) * 16384 + k * 32), 16, 1024);
                                            ptx::umma_bf16(d, adesc, bdesc, idesc, acc);
                                        }
                                    }
                        ptx::umma_commit(&bars->mma_done);
                    }
                    __syncwarp();
                }
    } else {
        // ===================== compute warps =====================
        const int q = warp;                                    // TMEM lane quadrant
        const bool up = lane >= 16;                            // this thread keeps image rows 3..5 (else 0..2)
        const int c = 16 * q + (lane & 15);                    // its channel
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int row0 = up ? 3 : 0;
        uint32_t mma_phase = 0;
        const float g1 = p.gamma[0][c], be1 = p.beta[0][c], g2 = p.gamma[1][c], be2 = p.beta[1][c], g3 = p.gamma[2][c], be3 = p.beta[2][c];
        const float cb1 = p.conv_b[0][c], cb2 = p.conv_b[1][c];

        // write relu(v) as bf16 hi / lo into the three shifted, padded, swizzled operand images
        auto store_operand = [&](const float* a) {
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const int h = row0 + j / kW, w = j % kW;
                __nv_bfloat16 hi, lo;
                split_bf16(a[j], hi, lo);
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int col = w + 1 - s;                 // column in the copy shifted left by s pixels
                    if (col < 0) continue;
                    const int P = (h + 1) * 8 + col;
                    const uint32_t off = (uint32_t)(P * 128 + ((((c >> 3) ^ (P & 7)) << 4) | ((c & 7) << 1)));
                    *reinterpret_cast<__nv_bfloat16*>(smem_act + (s * 2 + 0) * kImgBytes + off) = hi;
                    *reinterpret_cast<__nv_bfloat16*>(smem_act + (s * 2 + 1) * kImgBytes + off) = lo;
                }
            }
        };
        // accumulator of a convolution -> this thread's 18 pixels (hi + lo weight rows added)
        auto load_acc = [&](uint32_t dcol, float* out) {
            float v[kNWin];
            ptx::tmem_ld16(tb + lane_addr + dcol, v);
            ptx::tmem_ld16(tb + lane_addr + dcol + 16, v + 16);
            ptx::tmem_ld16(tb + lane_addr + dcol + 32, v + 32);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const int lo_col = (j / kW) * 8 + (j % kW), hi_col = lo_col + 24;        // rows 0..2 / rows 3..5
                const float send = up ? v[lo_col] : v[hi_col];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
                const float mine = up ? v[hi_col] : v[lo_col];
                out[j] = mine + recv;
            }
        };
        auto gidx = [&](int n, int j) { return ((size_t)n * kHW + (row0 * kW + j)) * kC + c; };
        auto split_store = [&](__nv_bfloat16* dst, int n, const float* a) {      // split tensor [B][H][2][W][C]
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const int h = row0 + j / kW, w = j % kW;
                __nv_bfloat16 hi, lo;
                split_bf16(a[j], hi, lo);
                dst[split_index(n, h, 0, w, c, kH, kW, kC)] = hi;
                dst[split_index(n, h, 1, w, c, kH, kW, kC)] = lo;
            }
        };

        for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
            float y[kPx], xi[kPx], kk[S > 1 ? S - 1 : 1][kPx];
#pragma unroll
            for (int j = 0; j < kPx; ++j) y[j] = p.x[gidx(n, j)];
            for (int step = 0; step < p.N; ++step) {
                const float dt = p.dt[step];
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const float ti = p.ts[step * MSB_MAX_STAGES + i];
                    char* slot = p.tape ? p.tape + (size_t)(step * S + i) * 5 * p.slot_stride : nullptr;
                    if (i == 0) {
#pragma unroll
                        for (int j = 0; j < kPx; ++j) xi[j] = y[j];
                    }
                    float a[kPx];
                    // ---- GN1 + ReLU -> operand of conv1 ----
                    {
                        float mean, rstd;
                        group_stats(xi, p.eps, mean, rstd);
#pragma unroll
                        for (int j = 0; j < kPx; ++j) {
                            const float v = (xi[j] - mean) * rstd * g1 + be1;
                            a[j] = v > 0.f ? v : 0.f;
                        }
                    }
                    if (slot) {
#pragma unroll
                        for (int j = 0; j < kPx; ++j) reinterpret_cast<float*>(slot)[gidx(n, j)] = xi[j];
                        split_store(reinterpret_cast<__nv_bfloat16*>(slot + 3 * p.slot_stride), n, a);
                    }
                    store_operand(a);
                    fence_async_smem();
                    ptx::tc_fence_before();
                    bar_sync(1, kThreads);                     // -> MMA warp: conv1
                    ptx::mbar_wait(&bars->mma_done, mma_phase); mma_phase ^= 1;
                    ptx::tc_fence_after();
                    // ---- P1 = conv1 + bias + t * tapmap;  GN2 + ReLU -> operand of conv2 ----
                    float pv[kPx];
                    load_acc(kD1Col, pv);
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        pv[j] = __fadd_rn(pv[j], cb1);
                        pv[j] = __fadd_rn(pv[j], __fmul_rn(ti, smem_map[(row0 * kW + j) * kC + c]));
                    }
                    {
                        float mean, rstd;
                        group_stats(pv, p.eps, mean, rstd);
#pragma unroll
                        for (int j = 0; j < kPx; ++j) {
                            const float v = (pv[j] - mean) * rstd * g2 + be2;
                            a[j] = v > 0.f ? v : 0.f;
                        }
                    }
                    if (slot) {
#pragma unroll
                        for (int j = 0; j < kPx; ++j) reinterpret_cast<float*>(slot + p.slot_stride)[gidx(n, j)] = pv[j];
                        split_store(reinterpret_cast<__nv_bfloat16*>(slot + 4 * p.slot_stride), n, a);
                    }
                    store_operand(a);
                    fence_async_smem();
                    ptx::tc_fence_before();
                    bar_sync(1, kThreads);                     // -> MMA warp: conv2
                    ptx::mbar_wait(&bars->mma_done, mma_phase); mma_phase ^= 1;
                    ptx::tc_fence_after();
                    // ---- P2 = conv2 + bias + t * tapmap;  k_i = GN3(P2);  Runge-Kutta combination ----
                    load_acc(kD2Col, pv);
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        pv[j] = __fadd_rn(pv[j], cb2);
                        pv[j] = __fadd_rn(pv[j], __fmul_rn(ti, smem_map[kMapFloats + (row0 * kW + j) * kC + c]));
                    }
                    if (slot) {
#pragma unroll
                        for (int j = 0; j < kPx; ++j) reinterpret_cast<float*>(slot + 2 * p.slot_stride)[gidx(n, j)] = pv[j];
                    }
                    float mean, rstd;
                    group_stats(pv, p.eps, mean, rstd);
                    // the same sums, in the same order, as epilogue_apply(): s = sum_j k_j coef_j + k_i coef_v; out = y + s dt
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        const float kv = (pv[j] - mean) * rstd * g3 + be3;
                        float s;
                        if (i < S - 1) {
                            s = 0.f;
                            bool first = true;
#pragma unroll
                            for (int m = 0; m < i; ++m) {
                                const float t = __fmul_rn(kk[m][j], p.w[(i + 1) * MSB_MAX_STAGES + m]);
                                s = first ? t : __fadd_rn(s, t);
                                first = false;
                            }
                            const float t = __fmul_rn(kv, p.w[(i + 1) * MSB_MAX_STAGES + i]);
                            s = first ? t : __fadd_rn(s, t);
                            kk[i][j] = kv;
                            xi[j] = __fadd_rn(y[j], __fmul_rn(s, dt));
                        } else {
                            s = 0.f;
                            bool first = true;
#pragma unroll
                            for (int m = 0; m < S - 1; ++m) {
                                const float t = __fmul_rn(kk[m][j], p.b[m]);
                                s = first ? t : __fadd_rn(s, t);
                                first = false;
                            }
                            const float t = __fmul_rn(kv, p.b[S - 1]);
                            s = first ? t : __fadd_rn(s, t);
                            y[j] = __fadd_rn(y[j], __fmul_rn(s, dt));
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kPx; ++j) p.y_out[gidx(n, j)] = y[j];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kComputeWarps) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, kTmemCols); }
}

}  // namespace

bool mnist_fused_supported(const MsbOdeDesc* d, const MsbMnistParams* mp) {
    if (!tune_get(TUNE_MNIST_FUSED)) return false;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    return major == 10 && d->channels == kC && d->height == kH && d->width == kW && mp && mp->groups == 32 && d->stages >= 1 &&
           d->stages <= MSB_MAX_STAGES && d->n_steps <= kMaxStepsFused && d->n_solvers <= 1;
}
size_t mnist_fused_workspace_bytes() { return align_up((size_t)128 * 288 * 4) + align_up((size_t)kW2Bytes) + 2 * align_up((size_t)kMapFloats * 4); }

// stage time t_n + c_i dt with the reference's two roundings (`_get_t`), as in odeblock.cu
static float stage_time(const MsbOdeDesc* d, int n, int i) {
    const float t0 = d->time_grid[n];
    const float dt = d->time_grid[n + 1] - t0;
    volatile float cdt = d->c[i] * dt;
    return (i == 0) ? t0 : t0 + cdt;
}

int launch_mnist_fused_forward(const MsbOdeDesc* d, const float* x, const MsbMnistParams* mp, float* y_out, void* workspace,
                               void* tape, size_t slot_stride, cudaStream_t st) {
    Carver cv(workspace, mnist_fused_workspace_bytes());
    void* w1img = cv.take<char>((size_t)128 * 288 * 4);
    __nv_bfloat16* w2t = cv.take<__nv_bfloat16>(kW2Bytes);
    float* tm[2] = {cv.take<float>(kMapFloats * 4), cv.take<float>(kMapFloats * 4)};
    launch_pack_w_tct_ex(mp->conv_w[0], w1img, 0, kC + 1, 1, st);
    launch_pack_w_tc_ex(mp->conv_w[1], w2t, kC, 0, kC + 1, 1, st);
    for (int k = 0; k < 2; ++k) launch_time_tapmap(mp->conv_w[k], tm[k], kH, kW, kC, st);
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.y_out = y_out; p.w1_tmem_img = (const uint4*)w1img; p.w2_tiles = w2t;
    for (int k = 0; k < 2; ++k) { p.tapmap[k] = tm[k]; p.conv_b[k] = mp->conv_b[k]; }
    for (int k = 0; k < 3; ++k) { p.gamma[k] = mp->norm_w[k]; p.beta[k] = mp->norm_b[k]; }
    p.eps = mp->eps; p.B = d->batch; p.S = d->stages; p.N = d->n_steps;
    memcpy(p.b, d->b, sizeof(p.b)); memcpy(p.w, d->w, sizeof(p.w));
    for (int n = 0; n < d->n_steps; ++n) {
        p.dt[n] = d->time_grid[n + 1] - d->time_grid[n];
        for (int i = 0; i < d->stages; ++i) p.ts[n * MSB_MAX_STAGES + i] = stage_time(d, n, i);
    }
    p.tape = (char*)tape; p.slot_stride = slot_stride;
    const size_t smem = (size_t)kW2Bytes + kActBytes + 2 * kMapFloats * 4 + sizeof(FBars) + 1024;
    const int grid = std::min(d->batch, num_sms());
    void (*kern)(const FusedParams) = nullptr;
    switch (d->stages) {
        case 1: kern = mnist_fused_fwd_kernel<1>; break;
        case 2: kern = mnist_fused_fwd_kernel<2>; break;
        case 3: kern = mnist_fused_fwd_kernel<3>; break;
        default: kern = mnist_fused_fwd_kernel<4>; break;
    }
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(mnist_fused)"))
        return -1;
    kern<<<grid, kThreads, smem, st>>>(p);
    count_launch();
    return check_cuda(cudaGetLastError(), "mnist fused forward launch");
}

}  // namespace msb
