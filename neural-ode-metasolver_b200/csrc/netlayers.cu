// Kernels of the non-ODE remainder of premetanode10:
//   * stem: 3x3 convolution 3 -> C channels + activation        (cifar10/layers.py:411-413)
//   * strided residual block (PreBasicBlock, stride 2, 1x1 stride-2 shortcut; layers.py:54-81): its three
//     convolutions run on the SAME tcgen05 / SIMT convolution engines as the ODE blocks after a
//     space-to-depth re-indexing, so the only new device code is the re-indexing itself:
//         s2d:  x (B,H,W,Ci) -> two "phase-row" operands T_p (B,H/2,W/2,2Ci), p = row parity, channels =
//               (column parity q, c), plus the shortcut operand T_sc = x[::2, ::2] zero-padded to 2Ci channels
//         a stride-2 3x3 convolution of x  ==  conv3x3(T_0, W'_0) + conv3x3(T_1, W'_1)   (stride 1, 2Ci -> Co)
//               with W'_p[co][q*Ci+ci][tr][ts] = W[co][ci][r][s] where input row 2i+r-1 = 2(i+tr-1)+p
//         d2s:  the adjoint re-indexing for the input gradient.
// All tensors fp32 NHWC / bf16 hi-lo split as everywhere else in the library.
#include "msb_internal.h"

namespace msb {

// ---------------------------------------------------------------------------------------------
// stem forward:  y = act(conv3x3(x, w)),  dact = act'(conv3x3(x, w)).   x: (B,H,W,3), w: (C,3,3,3) OIHW
// thread = 4 consecutive pixels of one image row x 16 output channels (each weight read from shared
// memory feeds 4 FMAs); weights transposed to [27][C] in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kStemPX = 4;
__global__ void __launch_bounds__(128) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int act,
                                                       float* __restrict__ y, float* __restrict__ dact, int B, int H, int W,
                                                       int C) {
    extern __shared__ float sw[];                      // [27][C]
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
        const int co = i % C, k = i / C;               // k = ci*9 + r*3 + s  (OIHW order inside one filter)
        sw[i] = w[(size_t)co * 27 + k];
    }
    __syncthreads();
    const int groups = C / 16;
    const int wgroups = (W + kStemPX - 1) / kStemPX;
    const long long total = (long long)B * H * wgroups * groups;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cg = (int)(t % groups);
    long long pg = t / groups;
    const int w0 = (int)(pg % wgroups) * kStemPX; pg /= wgroups;
    const int h = (int)(pg % H);
    const long long n = pg / H;
    float patch[3][kStemPX + 2][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < kStemPX + 2; ++cc) {
            const int ih = h + r - 1, iw = w0 + cc - 1;
            const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
            const float* p = x + (((size_t)n * H + (ok ? ih : 0)) * W + (ok ? iw : 0)) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) patch[r][cc][ci] = ok ? p[ci] : 0.f;
        }
    float acc[kStemPX][16];
#pragma unroll
    for (int px = 0; px < kStemPX; ++px)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[px][j] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const float4* wr = reinterpret_cast<const float4*>(sw + (ci * 9 + r * 3 + s) * C + cg * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = wr[q];
#pragma unroll
                    for (int px = 0; px < kStemPX; ++px) {
                        const float xv = patch[r][px + s][ci];
                        acc[px][4 * q + 0] = fmaf(xv, wv.x, acc[px][4 * q + 0]);
                        acc[px][4 * q + 1] = fmaf(xv, wv.y, acc[px][4 * q + 1]);
                        acc[px][4 * q + 2] = fmaf(xv, wv.z, acc[px][4 * q + 2]);
                        acc[px][4 * q + 3] = fmaf(xv, wv.w, acc[px][4 * q + 3]);
                    }
                }
            }
#pragma unroll
    for (int px = 0; px < kStemPX; ++px) {
        if (w0 + px >= W) break;
        const size_t pix = ((size_t)n * H + h) * W + w0 + px;
        float a[16], d[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) act_both(act, acc[px][j], a[j], d[j]);
        // 256-bit stores: a lane writes whole 32-byte sectors (four 16-byte stores per lane at a 64-byte lane stride left every
        // sector of a warp-wide store half written)
        stg256(y + pix * C + cg * 16, a);
        stg256(y + pix * C + cg * 16 + 8, a + 8);
        if (dact) {
            stg256(dact + pix * C + cg * 16, d);
            stg256(dact + pix * C + cg * 16 + 8, d + 8);
        }
    }
}

int launch_stem_fwd(const float* x, const float* w, int act, float* y, float* dact, int B, int H, int W, int C,
                    cudaStream_t st) {
    if (C % 16 || C > 256) { set_error("stem: output channels must be a multiple of 16, <= 256 (got %d)", C); return -1; }
    const long long threads = (long long)B * H * ((W + kStemPX - 1) / kStemPX) * (C / 16);
    stem_fwd_kernel<<<(unsigned)((threads + 127) / 128), 128, 27 * C * sizeof(float), st>>>(x, w, act, y, dact, B, H, W, C);
    count_launch();
    return check_cuda(cudaGetLastError(), "stem forward launch");
}

// ---------------------------------------------------------------------------------------------
// stem weight gradient:  partial[block][k][co] = sum over the block's pixels of gpre[p][co] * patch[p][k],
// gpre = gy * dact.  Thread = 4 output channels x 7 taps (27 padded to 28) x one quarter of the staged
// pixels; the four pixel quarters are summed through shared memory in a fixed order.  Deterministic: fixed
// pixel order per block + fixed-order reduction over blocks.
// ---------------------------------------------------------------------------------------------
constexpr int kStemPix = 64;   // pixels staged per iteration (45 KB of static shared memory)
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ dact,
                                                         const float* __restrict__ x, float* __restrict__ partial, int B,
                                                         int H, int W, int C, long long pix_per_block, int wsh, int hsh) {
    __shared__ __align__(16) float sg[kStemPix][64];
    __shared__ __align__(16) float sx[kStemPix][32];       // [pixel][k-group 0..3][8]: 7 taps + 1 pad per group
    __shared__ float sred[3][64][28];
    const int co4 = threadIdx.x & 15, kg = (threadIdx.x >> 4) & 3, ps = threadIdx.x >> 6;
    const long long P = (long long)B * H * W;
    const long long p_beg = (long long)blockIdx.x * pix_per_block;
    const long long p_end = p_beg + pix_per_block < P ? p_beg + pix_per_block : P;
    for (int cb = 0; cb < C; cb += 64) {
        float acc[4][7];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
        // global loads of chunk i+1 are issued before the FMAs of chunk i (register prefetch): the staging loop was
        // latency-bound (load -> shared -> barrier -> compute -> barrier, nothing in flight during the compute)
        constexpr int NG4 = kStemPix * 16 / 256, NX = kStemPix * 32 / 256;
        float4 rg[NG4];
        float rx[NX];
        auto load_chunk = [&](long long p0) {
#pragma unroll
            for (int t = 0; t < NG4; ++t) {                                    // gpre: kStemPix pixels x 16 float4
                const int i = threadIdx.x + t * 256;
                const int pl = i >> 4, c4 = i & 15;
                const long long pix = p0 + pl;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pix < p_end) {
                    const size_t idx = (size_t)pix * C + cb + c4 * 4;
                    const float4 g = *reinterpret_cast<const float4*>(gy + idx);
                    const float4 d = *reinterpret_cast<const float4*>(dact + idx);
                    v = make_float4(g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w);
                }
                rg[t] = v;
            }
#pragma unroll
            for (int t = 0; t < NX; ++t) {                                     // input patches
                const int i = threadIdx.x + t * 256;
                const int pl = i >> 5, slot = i & 31;
                const int k = (slot >> 3) * 7 + (slot & 7);
                const long long pix = p0 + pl;
                float v = 0.f;
                if (pix < p_end && (slot & 7) < 7 && k < 27) {
                    const int ci = k / 9, r = (k % 9) / 3, s = k % 3;
                    const unsigned upix = (unsigned)pix;                 // B*H*W < 2^31 (checked by the launcher): 32-bit divides
                    int wq, h;
                    long long n;
                    if (wsh >= 0 && hsh >= 0) {                           // power-of-two image: shifts instead of three divisions
                        wq = (int)(upix & (unsigned)(W - 1));             // per element (the staging loop was issue-bound on them)
                        const unsigned rest = upix >> wsh;
                        h = (int)(rest & (unsigned)(H - 1));
                        n = rest >> hsh;
                    } else {
                        wq = (int)(upix % (unsigned)W);
                        const unsigned rest = upix / (unsigned)W;
                        h = (int)(rest % (unsigned)H);
                        n = rest / (unsigned)H;
                    }
                    const int ih = h + r - 1, iw = wq + s - 1;
                    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((size_t)n * H + ih) * W + iw) * 3 + ci];
                }
                rx[t] = v;
            }
        };
        if (p_beg < p_end) load_chunk(p_beg);
        for (long long p0 = p_beg; p0 < p_end; p0 += kStemPix) {
#pragma unroll
            for (int t = 0; t < NG4; ++t) {
                const int i = threadIdx.x + t * 256;
                *reinterpret_cast<float4*>(&sg[i >> 4][(i & 15) * 4]) = rg[t];
            }
#pragma unroll
            for (int t = 0; t < NX; ++t) {
                const int i = threadIdx.x + t * 256;
                sx[i >> 5][i & 31] = rx[t];
            }
            if (p0 + kStemPix < p_end) load_chunk(p0 + kStemPix);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kStemPix / 4; ++q) {
                const int pl = ps * (kStemPix / 4) + q;
                const float4 g = *reinterpret_cast<const float4*>(&sg[pl][co4 * 4]);
                const float4 xa = *reinterpret_cast<const float4*>(&sx[pl][kg * 8]);
                const float4 xb = *reinterpret_cast<const float4*>(&sx[pl][kg * 8 + 4]);
                const float gg[4] = {g.x, g.y, g.z, g.w};
                const float xx[7] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 7; ++j) acc[i][j] = fmaf(gg[i], xx[j], acc[i][j]);
            }
            __syncthreads();
        }
        // sum the four pixel quarters in the fixed order 0 + 1 + 2 + 3
        if (ps > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 7; ++j) sred[ps - 1][co4 * 4 + i][kg * 7 + j] = acc[i][j];
        }
        __syncthreads();
        if (ps == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int k = kg * 7 + j, co = co4 * 4 + i;
                    const float v = ((acc[i][j] + sred[0][co][k]) + sred[1][co][k]) + sred[2][co][k];
                    if (k < 27) partial[((size_t)blockIdx.x * 27 + k) * C + cb + co] = v;
                }
        }
        __syncthreads();
    }
}

// grad_w[co][k] = sum_b partial[b][k][co]        (OIHW: k = ci*9 + r*3 + s)
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ gw, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * C) return;
    const int k = i % 27, co = i / 27;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += partial[((size_t)b * 27 + k) * C + co];
    gw[i] = acc;
}

int stem_wgrad_blocks() { return num_sms() * 2; }

int launch_stem_wgrad(const float* gy, const float* dact, const float* x, float* partial, float* gw, int B, int H, int W,
                      int C, cudaStream_t st) {
    if (C % 64) { set_error("stem wgrad: output channels must be a multiple of 64 (got %d)", C); return -1; }
    if ((long long)B * H * W >= (1LL << 31)) { set_error("stem wgrad: more than 2^31 pixels"); return -1; }
    const int nb = stem_wgrad_blocks();
    const long long P = (long long)B * H * W;
    long long per = (P + nb - 1) / nb;
    per = (per + kStemPix - 1) / kStemPix * kStemPix;
    auto log2_exact = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    stem_wgrad_kernel<<<nb, 256, 0, st>>>(gy, dact, x, partial, B, H, W, C, per, log2_exact(W), log2_exact(H));
    stem_wgrad_reduce_kernel<<<(27 * C + 127) / 128, 128, 0, st>>>(partial, nb, gw, C);
    count_launch(2);
    return check_cuda(cudaGetLastError(), "stem wgrad launch");
}

// ---------------------------------------------------------------------------------------------
// stem input gradient:  gx[p][ci] = sum_{r,s,co} gpre[p - (r-1, s-1)][co] * w[co][ci][r][s]
// Block = 8 x 32 pixel tile of one image; gpre staged through shared memory 16 channels at a time.
// ---------------------------------------------------------------------------------------------
constexpr int kDgTH = 8, kDgTW = 32, kDgCo = 16, kDgStride = 20;   // pixel stride 20 floats: conflict-free LDS.128
__global__ void __launch_bounds__(256) stem_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ dact,
                                                         const float* __restrict__ w, float* __restrict__ gx, int B, int H,
                                                         int W, int C) {
    __shared__ __align__(16) float sg[(kDgTH + 2) * (kDgTW + 2) * kDgStride];
    __shared__ __align__(16) float swt[9 * kDgCo * 4];                   // [tap][co][ci (3, padded to 4)]
    const int tiles_w = (W + kDgTW - 1) / kDgTW, tiles_h = (H + kDgTH - 1) / kDgTH;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h; b /= tiles_h;
    const int n = b;
    const int lr = threadIdx.x / kDgTW, lc = threadIdx.x % kDgTW;
    const int h = th * kDgTH + lr, wq = tw * kDgTW + lc;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < C; c0 += kDgCo) {
        for (int i = threadIdx.x; i < (kDgTH + 2) * (kDgTW + 2) * (kDgCo / 4); i += 256) {
            const int q = i % (kDgCo / 4), pl = i / (kDgCo / 4);
            const int pr = pl / (kDgTW + 2), pc = pl % (kDgTW + 2);
            const int ih = th * kDgTH + pr - 1, iw = tw * kDgTW + pc - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
                const size_t idx = (((size_t)n * H + ih) * W + iw) * C + c0 + q * 4;
                const float4 g = *reinterpret_cast<const float4*>(gy + idx);
                const float4 d = *reinterpret_cast<const float4*>(dact + idx);
                v = make_float4(g.x * d.x, g.y * d.y, g.z * d.z, g.w * d.w);
            }
            *reinterpret_cast<float4*>(sg + pl * kDgStride + q * 4) = v;
        }
        for (int i = threadIdx.x; i < 9 * kDgCo * 4; i += 256) {
            const int ci = i & 3, co = (i >> 2) % kDgCo, tap = i / (4 * kDgCo);
            swt[i] = ci < 3 ? w[((size_t)(c0 + co) * 3 + ci) * 9 + tap] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                // source pixel (h - (r-1), w - (s-1)) -> tile coordinates (+1 halo)
                const float* g = sg + ((lr + 2 - r) * (kDgTW + 2) + (lc + 2 - s)) * kDgStride;
                const float* wt = swt + (r * 3 + s) * kDgCo * 4;
#pragma unroll
                for (int q = 0; q < kDgCo / 4; ++q) {
                    const float4 gv = *reinterpret_cast<const float4*>(g + q * 4);
                    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 wv = *reinterpret_cast<const float4*>(wt + (q * 4 + j) * 4);
                        acc[0] = fmaf(gg[j], wv.x, acc[0]);
                        acc[1] = fmaf(gg[j], wv.y, acc[1]);
                        acc[2] = fmaf(gg[j], wv.z, acc[2]);
                    }
                }
            }
        __syncthreads();
    }
    if (h < H && wq < W) {
        float* o = gx + (((size_t)n * H + h) * W + wq) * 3;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    }
}

int launch_stem_dgrad(const float* gy, const float* dact, const float* w, float* gx, int B, int H, int W, int C,
                      cudaStream_t st) {
    if (C % kDgCo) { set_error("stem dgrad: output channels must be a multiple of %d (got %d)", kDgCo, C); return -1; }
    const int tiles = ((W + kDgTW - 1) / kDgTW) * ((H + kDgTH - 1) / kDgTH);
    stem_dgrad_kernel<<<(unsigned)(B * tiles), 256, 0, st>>>(gy, dact, w, gx, B, H, W, C);
    count_launch();
    return check_cuda(cudaGetLastError(), "stem dgrad launch");
}

// ---------------------------------------------------------------------------------------------
// space-to-depth prologue of the strided residual block.  x: (B,H,W,Ci) fp32.  Ho = H/2, Wo = W/2, Cc = 2*Ci.
//   T[p] (split [B][Ho][2][Wo][Cc]) : channel q*Ci + c of pixel (i,j) = hi/lo(act(x[2i+p][2j+q][c]))
//   Tsc  (same shape)               : channel c < Ci = hi/lo(x[2i][2j][c]) (no activation), channels >= Ci = 0
//   G0   (fp32, shape of x, optional) = act'(x)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s2d_act_split_kernel(const float* __restrict__ x, int act,
                                                            __nv_bfloat16* __restrict__ T0, __nv_bfloat16* __restrict__ T1,
                                                            __nv_bfloat16* __restrict__ Tsc, float* __restrict__ G0,
                                                            size_t n_vec, int H, int W, int Ci) {
    const int cv = Ci >> 2;
    const int Ho = H >> 1, Wo = W >> 1, Cc = 2 * Ci;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % cv);
        size_t pix = i / cv;
        const int wq = (int)(pix % W); pix /= W;
        const int h = (int)(pix % H);
        const int n = (int)(pix / H);
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float xv[4] = {v.x, v.y, v.z, v.w};
        float a[4], d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) act_both(act, xv[k], a[k], d[k]);
        if (G0) reinterpret_cast<float4*>(G0)[i] = make_float4(d[0], d[1], d[2], d[3]);
        const int p = h & 1, q = wq & 1, io = h >> 1, jo = wq >> 1;
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16(a[k], hi[k], lo[k]);
        __nv_bfloat16* T = p ? T1 : T0;
        const size_t o_hi = split_index(n, io, 0, jo, q * Ci + c4 * 4, Ho, Wo, Cc);
        const size_t o_lo = split_index(n, io, 1, jo, q * Ci + c4 * 4, Ho, Wo, Cc);
        *reinterpret_cast<uint2*>(T + o_hi) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(T + o_lo) = *reinterpret_cast<uint2*>(lo);
        if (p == 0 && q == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(xv[k], hi[k], lo[k]);
            const size_t s_hi = split_index(n, io, 0, jo, c4 * 4, Ho, Wo, Cc);
            const size_t s_lo = split_index(n, io, 1, jo, c4 * 4, Ho, Wo, Cc);
            *reinterpret_cast<uint2*>(Tsc + s_hi) = *reinterpret_cast<uint2*>(hi);
            *reinterpret_cast<uint2*>(Tsc + s_lo) = *reinterpret_cast<uint2*>(lo);
            const uint2 z = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(Tsc + s_hi + Ci) = z;
            *reinterpret_cast<uint2*>(Tsc + s_lo + Ci) = z;
        }
    }
}

void launch_s2d_act_split(const float* x, int act, __nv_bfloat16* T0, __nv_bfloat16* T1, __nv_bfloat16* Tsc, float* G0,
                          int B, int H, int W, int Ci, cudaStream_t st) {
    const size_t n_vec = (size_t)B * H * W * Ci / 4;
    int blocks = (int)std::min<size_t>((n_vec + 255) / 256, (size_t)num_sms() * 8);
    if (blocks < 1) blocks = 1;
    s2d_act_split_kernel<<<blocks, 256, 0, st>>>(x, act, T0, T1, Tsc, G0, n_vec, H, W, Ci);
    count_launch();
}

// adjoint:  gx[2i+p][2j+q][c] = gT_p[i][j][q*Ci+c] * G0[2i+p][2j+q][c]  (+ gTsc[i][j][c] when p = q = 0)
__global__ void __launch_bounds__(256) d2s_grad_kernel(const float* __restrict__ gT0, const float* __restrict__ gT1,
                                                       const float* __restrict__ gTsc, const float* __restrict__ G0,
                                                       float* __restrict__ gx, size_t n_vec, int H, int W, int Ci) {
    const int cv = Ci >> 2;
    const int Ho = H >> 1, Wo = W >> 1, Cc = 2 * Ci;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % cv);
        size_t pix = i / cv;
        const int wq = (int)(pix % W); pix /= W;
        const int h = (int)(pix % H);
        const int n = (int)(pix / H);
        const int p = h & 1, q = wq & 1, io = h >> 1, jo = wq >> 1;
        const size_t src = (((size_t)n * Ho + io) * Wo + jo) * Cc;
        const float4 g = *reinterpret_cast<const float4*>((p ? gT1 : gT0) + src + q * Ci + c4 * 4);
        const float4 d = reinterpret_cast<const float4*>(G0)[i];
        float4 o = make_float4(__fmul_rn(g.x, d.x), __fmul_rn(g.y, d.y), __fmul_rn(g.z, d.z), __fmul_rn(g.w, d.w));
        if (p == 0 && q == 0) {
            const float4 s = *reinterpret_cast<const float4*>(gTsc + src + c4 * 4);
            o.x = __fadd_rn(o.x, s.x); o.y = __fadd_rn(o.y, s.y); o.z = __fadd_rn(o.z, s.z); o.w = __fadd_rn(o.w, s.w);
        }
        reinterpret_cast<float4*>(gx)[i] = o;
    }
}

void launch_d2s_grad(const float* gT0, const float* gT1, const float* gTsc, const float* G0, float* gx, int B, int H, int W,
                     int Ci, cudaStream_t st) {
    const size_t n_vec = (size_t)B * H * W * Ci / 4;
    int blocks = (int)std::min<size_t>((n_vec + 255) / 256, (size_t)num_sms() * 8);
    if (blocks < 1) blocks = 1;
    d2s_grad_kernel<<<blocks, 256, 0, st>>>(gT0, gT1, gTsc, G0, gx, n_vec, H, W, Ci);
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// weights of the strided block in space-to-depth form (fp32 OIHW, Co x Cc x 3 x 3, Cc = 2*Ci = Co):
//   Wd[p][co][q*Ci+ci][tr][ts] = w1[co][ci][r][s]  with (p, tr) <-> r:  r=0:(1,0)  r=1:(0,1)  r=2:(1,1), same for (q, ts) <-> s
//   Wd[2][co][ci][1][1]        = wsc[co][ci]  (ci < Ci), everything else 0
// and the adjoint gather of their gradients.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tap_of(int parity, int t) {       // original tap index r for (parity, t) or -1
    if (parity == 0) return t == 1 ? 1 : -1;
    return t == 0 ? 0 : (t == 1 ? 2 : -1);
}
__global__ void down_weights_build_kernel(const float* __restrict__ w1, const float* __restrict__ wsc,
                                          float* __restrict__ Wd, int Ci, int Co) {
    const int Cc = 2 * Ci;
    const int per = Co * Cc * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * per; i += gridDim.x * blockDim.x) {
        const int which = i / per;
        int j = i - which * per;
        const int ts = j % 3; j /= 3;
        const int tr = j % 3; j /= 3;
        const int cc = j % Cc;
        const int co = j / Cc;
        float v = 0.f;
        if (which < 2) {
            const int q = cc / Ci, ci = cc - q * Ci;
            const int r = tap_of(which, tr), s = tap_of(q, ts);
            if (r >= 0 && s >= 0) v = w1[(((size_t)co * Ci + ci) * 3 + r) * 3 + s];
        } else if (cc < Ci && tr == 1 && ts == 1) {
            v = wsc[(size_t)co * Ci + cc];
        }
        Wd[i] = v;
    }
}
__global__ void down_weights_gather_kernel(const float* __restrict__ gWd, float* __restrict__ gw1, float* __restrict__ gwsc,
                                           int Ci, int Co) {
    const int Cc = 2 * Ci;
    const size_t per = (size_t)Co * Cc * 9;
    const int n1 = Co * Ci * 9, n2 = Co * Ci;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            int j = i;
            const int s = j % 3; j /= 3;
            const int r = j % 3; j /= 3;
            const int ci = j % Ci;
            const int co = j / Ci;
            const int p = r == 1 ? 0 : 1, tr = r == 0 ? 0 : 1;
            const int q = s == 1 ? 0 : 1, ts = s == 0 ? 0 : 1;
            gw1[i] = gWd[(size_t)p * per + (((size_t)co * Cc + q * Ci + ci) * 3 + tr) * 3 + ts];
        } else {
            const int j = i - n1;
            const int ci = j % Ci, co = j / Ci;
            gwsc[j] = gWd[2 * per + (((size_t)co * Cc + ci) * 3 + 1) * 3 + 1];
        }
    }
}

void launch_down_weights_build(const float* w1, const float* wsc, float* Wd, int Ci, int Co, cudaStream_t st) {
    const int total = 3 * Co * 2 * Ci * 9;
    down_weights_build_kernel<<<(total + 255) / 256, 256, 0, st>>>(w1, wsc, Wd, Ci, Co);
    count_launch();
}
void launch_down_weights_gather(const float* gWd, float* gw1, float* gwsc, int Ci, int Co, cudaStream_t st) {
    const int total = Co * Ci * 10;
    down_weights_gather_kernel<<<(total + 255) / 256, 256, 0, st>>>(gWd, gw1, gwsc, Ci, Co);
    count_launch();
}

}  // namespace msb
