// tcgen05 engine: weight gradient of the 3x3 convolution as a split-K GEMM over pixels.
//
//   dW[c_out][tap][c_in] = sum_pixels gout[pixel][c_out] * in[pixel + tap][c_in]
//
// Both operands are read in place from split tensors [B][H][2][W][C] (channels contiguous), i.e.
// they are "MN-major" for this GEMM (K = pixels is the strided dimension):
//   A (M = 128) : gout.  C=64 : rows 0..63 = hi plane, 64..127 = lo plane of the 64 channels
//                              (two 64-element MN atoms, LBO = plane stride)
//                        C=128: the 128 channels of one plane (two 64-channel TMA boxes, LBO = box stride);
//                              hi and lo planes are separate MMAs
//   B (N = 192) : in, 3 vertical taps x 64 input channels: three MN atoms whose stride (LBO) is one
//                 image row of the halo tile, so one MMA covers r = 0,1,2
//   K = 16      : 16 consecutive pixels of one image row (two 8-pixel swizzle atoms, SBO = 1024 B)
// All four hi/lo products are accumulated in fp32 in one TMEM accumulator (128 lanes x 192 cols),
// which stays resident for ALL tiles of a CTA: the only global write is one 128x192 partial at the
// end (optionally accumulated onto the CTA's own slot from earlier launches, so a whole backward pass
// needs a single reduction per weight tensor).  CTA = (group, part): group = (horizontal tap s, 64-wide c_in chunk), part = slice of the
// pixel tiles.  Partials are summed in a fixed order by wgrad_reduce_kernel (deterministic).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc, 4..7 = final TMEM -> global.
#include <cuda.h>

#include "msb_internal.h"
#include "msb_ptx.cuh"

namespace msb {

int make_tmap_split5d(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h);

namespace {

constexpr int kThreads = 256;
constexpr int kStages = 2;
constexpr uint32_t kTmemCols = 256;

template <int C, int WIMG> struct WG {
    static constexpr int ROWS = 128 / WIMG;
    static constexpr int CO_CHUNKS = C / 64;
    static constexpr int CI_CHUNKS = C / 64;
    static constexpr int GROUPS = 3 * CI_CHUNKS;
    static constexpr int GO_CHUNK_BYTES = ROWS * 2 * WIMG * 128;          // one 64-channel box of gout
    static constexpr int GO_BYTES = CO_CHUNKS * GO_CHUNK_BYTES;
    static constexpr int IN_BYTES = (ROWS + 2) * 2 * WIMG * 128;          // halo box of in
    static constexpr int STAGE_BYTES = GO_BYTES + IN_BYTES;
    static constexpr int ROW_PAIR_BYTES = 2 * WIMG * 128;
    static constexpr int PLANE_BYTES = WIMG * 128;
    static constexpr int HALVES = (C == 64) ? 2 : 1;                      // partial slices written per CTA
};

struct __align__(8) WBarriers {
    uint64_t full[kStages], empty[kStages];
    uint64_t gempty[kStages];          // cluster form, rank 0: every CTA of the cluster has released the stage
    uint64_t done;
    uint32_t tmem_base;
};

// MC: the GROUPS CTAs that share a pixel slice -- (horizontal tap, c_in chunk) pairs -- form one thread-block cluster and
// the gout box, identical for all of them, is loaded ONCE by rank 0 and multicast into every CTA's stage.  ncu had the
// kernel at 91 % of the L2 slice throughput cap (983 MB of L2 -> SM traffic per C = 64 launch, 9.4 TB/s): it was
// L2-bound, and a third (C = 64) / half (C = 128) of that traffic was the same gout tile fetched by each group's CTA.
template <int C, int WIMG, bool MC>
__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_go, const __grid_constant__ CUtensorMap tmap_in,
                   float* __restrict__ partial, const int num_tiles, const int tiles_per_img, const int nparts,
                   const int accumulate_partial, const uint32_t backoff_ns) {
    using G = WG<C, WIMG>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    WBarriers* bars = reinterpret_cast<WBarriers*>(smem + kStages * G::STAGE_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // plain launch: blocks [group][part]; cluster launch: blocks [part][group] so that a cluster = the groups of one part
    const int group = MC ? (int)(blockIdx.x % G::GROUPS) : (int)(blockIdx.x / nparts);
    const int part = MC ? (int)(blockIdx.x / G::GROUPS) : (int)(blockIdx.x - group * nparts);
    constexpr uint16_t kAllCtas = (uint16_t)((1u << G::GROUPS) - 1);
    const int s = group % 3;
    const int ci_chunk = group / 3;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_go);
        ptx::prefetch_tmap(&tmap_in);
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(&bars->full[i], 1); ptx::mbar_init(&bars->empty[i], 1); ptx::mbar_init(&bars->gempty[i], G::GROUPS);
        }
        ptx::mbar_init(&bars->done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(&bars->tmem_base, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (MC) ptx::cluster_sync();         // every CTA's barriers exist before rank 0 multicasts into them
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    ptx::pdl_launch_dependents();
    ptx::pdl_wait();                    // operands (and the partial slots accumulated in place) are the predecessors' outputs

    if (warp == 0) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int tile = part; tile < num_tiles; tile += nparts) {
                const int n = tile / tiles_per_img;
                const int h0 = (tile - n * tiles_per_img) * G::ROWS;
                uint8_t* stage = smem + st * G::STAGE_BYTES;
                ptx::mbar_wait(&bars->empty[st], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&bars->full[st], G::STAGE_BYTES);
                if (MC) {
                    if (group == 0) {            // (group == cluster rank) one gout load for the whole cluster
                        ptx::mbar_wait(&bars->gempty[st], ph ^ 1);
                        for (int cc = 0; cc < G::CO_CHUNKS; ++cc)
                            ptx::tma_load_5d_multicast(stage + cc * G::GO_CHUNK_BYTES, &tmap_go, &bars->full[st], kAllCtas,
                                                       cc * 64, 0, 0, h0, n);
                    }
                } else
                for (int cc = 0; cc < G::CO_CHUNKS; ++cc)
                    ptx::tma_load_5d(stage + cc * G::GO_CHUNK_BYTES, &tmap_go, &bars->full[st], cc * 64, 0, 0, h0, n);
                ptx::tma_load_5d(stage + G::GO_BYTES, &tmap_in, &bars->full[st], ci_chunk * 64, s - 1, 0, h0 - 1, n);
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 192, 1, 1);
            constexpr uint32_t a_lbo = (C == 64) ? G::PLANE_BYTES : G::GO_CHUNK_BYTES;
            int st = 0; uint32_t ph = 0;
            uint32_t accumulate = 0;
            for (int tile = part; tile < num_tiles; tile += nparts) {
                ptx::mbar_wait(&bars->full[st], ph);
                ptx::tc_fence_after();
                const uint32_t go_base = ptx::smem_u32(smem + st * G::STAGE_BYTES);
                const uint32_t in_base = go_base + G::GO_BYTES;
                for (int rho = 0; rho < G::ROWS; ++rho)
                    for (int wg = 0; wg < WIMG / 16; ++wg) {
                        const uint32_t px_off = (uint32_t)wg * 16 * 128;
#pragma unroll
                        for (int pa = 0; pa < (C == 64 ? 1 : 2); ++pa) {
                            const uint64_t adesc = ptx::make_smem_desc_sw128(
                                go_base + rho * G::ROW_PAIR_BYTES + pa * G::PLANE_BYTES + px_off, a_lbo, 1024);
#pragma unroll
                            for (int pb = 0; pb < 2; ++pb) {
                                if (C != 64 && pa == 1 && pb == 1) continue;   // lo x lo (2^-18 relative) is dropped
                                const uint64_t bdesc = ptx::make_smem_desc_sw128(
                                    in_base + rho * G::ROW_PAIR_BYTES + pb * G::PLANE_BYTES + px_off, G::ROW_PAIR_BYTES, 1024);
                                ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, accumulate);
                                accumulate = 1;
                            }
                        }
                    }
                ptx::umma_commit(&bars->empty[st]);
                if (MC) ptx::umma_commit_multicast(&bars->gempty[st], (uint16_t)1);      // -> rank 0
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
            ptx::umma_commit(&bars->done);
        }
    } else if (warp >= 4) {
        // final epilogue: D[lane = gout channel (or hi/lo half x channel)][col = r*64 + ci] -> partial
        const int q = warp & 3;
        ptx::mbar_wait_backoff(&bars->done, 0, backoff_ns ? 8 * backoff_ns : 0);     // waits for the whole kernel
        ptx::tc_fence_after();
        const int row = q * 32 + lane;                       // accumulator row
        int slice, co;
        if (C == 64) { slice = part * 2 + (row >> 6); co = row & 63; }
        else { slice = part; co = row; }
        float* dst = partial + (size_t)slice * 9 * C * C;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < 192; cb += 16) {
            float v[16];
            ptx::tmem_ld16(t_addr + cb, v);
            ptx::tmem_ld_wait();
            const int r = cb >> 6;
            const int tap = r * 3 + s;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int ci = ci_chunk * 64 + (cb & 63) + j;
                float* q = dst + ((size_t)tap * C + ci) * C + co;
                *q = accumulate_partial ? *q + v[j] : v[j];        // CTA-private slot: deterministic
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (MC) ptx::cluster_sync();         // nobody leaves while a peer may still multicast into / signal this CTA
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int C, int WIMG>
int nparts_impl(ConvShape s) {
    using G = WG<C, WIMG>;
    const int num_tiles = s.B * (s.H / G::ROWS);
    int np = num_sms() / G::GROUPS;
    if (np > num_tiles) np = num_tiles;
    if (np < 1) np = 1;
    return np;
}

template <int C, int WIMG>
int launch_impl(const __nv_bfloat16* gout, const __nv_bfloat16* in, float* partial, int* nparts_out, int accumulate,
                ConvShape s, cudaStream_t st) {
    using G = WG<C, WIMG>;
    CUtensorMap tm_go, tm_in;
    if (make_tmap_split5d(&tm_go, gout, s.B, s.H, s.W, s.C, WIMG, G::ROWS)) return -1;
    if (make_tmap_split5d(&tm_in, in, s.B, s.H, s.W, s.C, WIMG, G::ROWS + 2)) return -1;
    const size_t smem = (size_t)kStages * G::STAGE_BYTES + sizeof(WBarriers) + 1024;
    const bool mc = tune_get(TUNE_WGRAD_MULTICAST) != 0;
    auto kern = mc ? wgrad3x3_tc_kernel<C, WIMG, true> : wgrad3x3_tc_kernel<C, WIMG, false>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(wgrad3x3_tc)"))
        return -1;
    const int tiles_per_img = s.H / G::ROWS;
    const int num_tiles = s.B * tiles_per_img;
    const int np = nparts_impl<C, WIMG>(s);
    cudaError_t le;
    if (mc) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(np * G::GROUPS)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = G::GROUPS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = tune_get(TUNE_PDL) ? 2 : 1;
        le = cudaLaunchKernelEx(&cfg, kern, tm_go, tm_in, partial, num_tiles, tiles_per_img, np, accumulate,
                                (uint32_t)tune_get(TUNE_WAIT_BACKOFF));
    } else {
        le = launch_maybe_pdl(kern, np * G::GROUPS, kThreads, smem, st, tm_go, tm_in, partial, num_tiles, tiles_per_img, np,
                              accumulate, (uint32_t)tune_get(TUNE_WAIT_BACKOFF));
    }
    count_launch();
    *nparts_out = np * G::HALVES;
    return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "wgrad3x3_tc launch");
}

}  // namespace

bool wgrad_tc_supported(ConvShape s) { return tc_shape_supported(s.C, s.H, s.W); }

int wgrad_tc_nparts(ConvShape s) {
    if (s.C == 64 && s.W == 32) return nparts_impl<64, 32>(s) * 2;
    if (s.C == 64 && s.W == 16) return nparts_impl<64, 16>(s) * 2;
    if (s.C == 128 && s.W == 32) return nparts_impl<128, 32>(s);
    return nparts_impl<128, 16>(s);
}

int launch_wgrad3x3_tc(const __nv_bfloat16* gout, const __nv_bfloat16* in, float* partial, int* nparts_out,
                       int accumulate, ConvShape s, cudaStream_t st) {
    if (!wgrad_tc_supported(s)) { set_error("tcgen05 wgrad: unsupported shape C=%d H=%d W=%d", s.C, s.H, s.W); return -1; }
    if (s.C == 64 && s.W == 32) return launch_impl<64, 32>(gout, in, partial, nparts_out, accumulate, s, st);
    if (s.C == 64 && s.W == 16) return launch_impl<64, 16>(gout, in, partial, nparts_out, accumulate, s, st);
    if (s.C == 128 && s.W == 32) return launch_impl<128, 32>(gout, in, partial, nparts_out, accumulate, s, st);
    return launch_impl<128, 16>(gout, in, partial, nparts_out, accumulate, s, st);
}

}  // namespace msb
