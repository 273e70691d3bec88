// tcgen05 engine: weight gradient of the 3x3 convolution as a split-K GEMM over pixels.
//
//   dW[c_out][tap][c_in] = sum_pixels gout[pixel][c_out] * in[pixel + tap][c_in]
//
// Both operands are read in place from split tensors [B][H][2][W][C] (channels contiguous), i.e.
// they are "MN-major" for this GEMM (K = pixels is the strided dimension):
//   C = 128:
//   A (M = 128) : gout, the 128 channels of one plane (two 64-channel TMA boxes, LBO = box stride);
//                 hi and lo planes are separate MMAs
//   B (N = 192) : in, 3 vertical taps x 64 input channels: three MN atoms whose stride (LBO) is one
//                 image row of the halo tile, so one MMA covers r = 0,1,2
//   K = 16      : 16 consecutive pixels of one image row (two 8-pixel swizzle atoms, SBO = 1024 B)
//   C = 64: roles swapped (`in` on M, two vertical taps per M = 128 operand; [g_hi | g_lo] on N) -- see the MMA loop.
// Three hi/lo products (lo x lo dropped; kept for the unpaired third tap at C = 64) are accumulated in fp32 in TMEM (128 lanes x 192 / 256 cols),
// which stays resident for ALL tiles of a CTA: the only global write is one 128x192 partial at the
// end (optionally accumulated onto the CTA's own slot from earlier launches, so a whole backward pass
// needs a single reduction per weight tensor).  CTA = (group, part): group = (horizontal tap s, 64-wide c_in chunk), part = slice of the
// pixel tiles.  Partials are summed in a fixed order by wgrad_reduce_kernel (deterministic).
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc, 4..7 = final TMEM -> global.
#include <cuda.h>

#include "msb_internal.h"
#include "msb_ptx.cuh"

#ifdef MSB_CONV_DEBUG
// instrumented build: clocks summed over CTAs -- [0] MMA warp waiting for a full stage, [1] its whole loop, [2] producer waiting
// for an empty stage, [3] kernel entry -> MMA loop start, [4] MMA loop end -> exit, [5] CTAs
static __device__ unsigned long long g_wgrad_dbg[8];
extern "C" int msb_debug_wgrad_read(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_wgrad_dbg, sizeof(g_wgrad_dbg)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long zero[8] = {0};
        if (cudaMemcpyToSymbol(g_wgrad_dbg, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

namespace msb {

int make_tmap_split5d(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h);

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 4;
constexpr uint32_t kTmemCols = 256;

// HT ("horizontal taps on N"; everything but the role-swapped C = 64 three-product option): a CTA owns one VERTICAL tap r and
// the N = 192 operand is the three HORIZONTAL taps -- three 64-channel atoms ONE PIXEL (LBO = 128 B) apart in a single staged
// copy of the image rows with a halo pixel each side.  tcgen05.mma reads a swizzled operand from any 128-byte row offset
// (scripts/probes/mma_rowshift_probe.cu, tests 2 and 4), so the staged `in` box is ROWS x (W + 2) pixels instead of
// (ROWS + 2) x W: 34 816 instead of 49 152 bytes per tile at C = 64 -- the kernel is bound by L2 -> SM operand traffic -- and
// three ring stages fit where two did.  Round 1 / HT = false: a CTA owns one horizontal tap s (its own shifted box), the three
// atoms are the vertical taps, one image row (LBO = a row pair) apart.
// MEASURED (option wgrad_htaps, default 1): identical results, L2 -> SM bytes 1021 -> 845 MB per C = 64 launch, the issuing
// warp's wait for full stages 29 % -> 12 % of its loop, MMA loop 163k -> 132k clocks per launch (scripts/diag_wgrad_waits.py);
// 142 -> 128 us per launch in the power-capped steady state (scripts/wgrad_htaps_ab.py).  C = 128 is neutral.  (The first
// measurement had it at 180 us: in that instantiation the compiler serialised the final accumulate-into-the-partial pass,
// see the comment there.)
template <int C, int WIMG, bool HT> struct WG {
    static constexpr int ROWS = 128 / WIMG;
    static constexpr int CO_CHUNKS = C / 64;
    static constexpr int CI_CHUNKS = C / 64;
    static constexpr int GROUPS = 3 * CI_CHUNKS;
    static constexpr int GO_CHUNK_BYTES = ROWS * 2 * WIMG * 128;          // one 64-channel box of gout
    static constexpr int GO_BYTES = CO_CHUNKS * GO_CHUNK_BYTES;
    static constexpr int ROW_PAIR_BYTES = 2 * WIMG * 128;                 // gout: [row][plane][pixel]
    static constexpr int PLANE_BYTES = WIMG * 128;
    static constexpr int IN_W = HT ? WIMG + 2 : WIMG;                     // staged pixels per image row of `in`
    static constexpr int IN_ROWS = HT ? ROWS : ROWS + 2;
    static constexpr int IN_PLANE_BYTES = IN_W * 128;
    static constexpr int IN_ROW_PAIR_BYTES = 2 * IN_PLANE_BYTES;
    static constexpr int IN_BYTES = IN_ROWS * IN_ROW_PAIR_BYTES;
    static_assert(IN_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned");
    static constexpr int STAGE_BYTES = GO_BYTES + IN_BYTES;
#ifdef MSB_WGRAD_STAGES_CAP
    static constexpr int STAGES = (225 * 1024 / STAGE_BYTES) > MSB_WGRAD_STAGES_CAP ? MSB_WGRAD_STAGES_CAP : (225 * 1024 / STAGE_BYTES);
#else
    static constexpr int STAGES = (225 * 1024 / STAGE_BYTES) > kMaxStages ? kMaxStages : (225 * 1024 / STAGE_BYTES);
#endif
    static_assert(STAGES >= 2, "operand ring");
    static constexpr int B_LBO = HT ? 128 : IN_ROW_PAIR_BYTES;            // stride between the three N atoms
    static constexpr int HALVES = (C == 64) ? 2 : 1;                      // partial slices written per CTA
};

struct __align__(8) WBarriers {
    uint64_t full[kMaxStages], empty[kMaxStages];
    uint64_t gempty[kMaxStages];       // cluster form, rank 0: every CTA of the cluster has released the stage
    uint64_t done;
    uint32_t tmem_base;
};

// MC: the GROUPS CTAs that share a pixel slice -- (horizontal tap, c_in chunk) pairs -- form one thread-block cluster and
// the gout box, identical for all of them, is loaded ONCE by rank 0 and multicast into every CTA's stage.  ncu had the
// kernel at 91 % of the L2 slice throughput cap (983 MB of L2 -> SM traffic per C = 64 launch, 9.4 TB/s): it was
// L2-bound, and a third (C = 64) / half (C = 128) of that traffic was the same gout tile fetched by each group's CTA.
template <int C, int WIMG, bool MC, bool P3, bool HT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_go, const __grid_constant__ CUtensorMap tmap_in,
                   float* __restrict__ partial, const int num_tiles, const int tiles_per_img, const int nparts,
                   const int accumulate_partial, const uint32_t backoff_ns, const int uniform_issue) {
    static_assert(!(HT && C == 64 && P3), "the role-swapped three-product form keeps the vertical taps on its operand");
    using G = WG<C, WIMG, HT>;
    constexpr int kStages = G::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    WBarriers* bars = reinterpret_cast<WBarriers*>(smem + kStages * G::STAGE_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
#ifdef MSB_CONV_DEBUG
    __shared__ long long dbg_entry, dbg_loop_end, dbg_epi_start;
    if (threadIdx.x == 0) dbg_entry = clock64();
#endif
    // plain launch: blocks [group][part]; cluster launch: blocks [part][group] so that a cluster = the groups of one part
    const int group = MC ? (int)(blockIdx.x % G::GROUPS) : (int)(blockIdx.x / nparts);
    const int part = MC ? (int)(blockIdx.x / G::GROUPS) : (int)(blockIdx.x - group * nparts);
    constexpr uint16_t kAllCtas = (uint16_t)((1u << G::GROUPS) - 1);
    const int s = group % 3;            // the tap index this CTA owns: horizontal (HT = false) or VERTICAL (HT = true)
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
#ifdef MSB_CONV_DEBUG
                { const long long t = clock64(); ptx::mbar_wait(&bars->empty[st], ph ^ 1); atomicAdd(&g_wgrad_dbg[2], (unsigned long long)(clock64() - t)); }
#else
                ptx::mbar_wait(&bars->empty[st], ph ^ 1);
#endif
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
                if (HT) ptx::tma_load_5d(stage + G::GO_BYTES, &tmap_in, &bars->full[st], ci_chunk * 64, -1, 0, h0 - 1 + s, n);
                else ptx::tma_load_5d(stage + G::GO_BYTES, &tmap_in, &bars->full[st], ci_chunk * 64, s - 1, 0, h0 - 1, n);
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && C == 64 && P3) {
            // C = 64, three hi/lo products (option wgrad64_products = 3; measured SLOWER than the four-product form, 161 vs
            // 143 us per launch: with both operands in shared memory the N = 128 / 64 MMAs need 128 - 192 B/clk of operand
            // reads, above the 128 B/clk shared-memory feed, while the N = 192 MMAs of the four-product form need 104 B/clk --
            // so it is not the default).  The roles are swapped against the C = 128 form: `in` rides on M, gout on N,
            // so that every row of an M = 128 operand needs the same N operand:
            //   D1[(r, c_in) : r = 0, 1][g_hi | g_lo]  += in_hi(rows rho, rho + 1) x [g_hi | g_lo]        (M = 128, N = 128)
            //   D1[(r, c_in)           ][g_hi       ]  += in_lo(rows rho, rho + 1) x  g_hi                 (M = 128, N = 64)
            //   D2[(plane, c_in) of r = 2][g_hi | g_lo] += [in_hi ; in_lo](row rho + 2) x [g_hi | g_lo]    (M = 128, N = 128)
            // 160 tensor clocks per 16 pixels instead of 192 (the third vertical tap has no partner tap and keeps its
            // lo x lo product).  Operands are still read in place: M / N atoms of 64 channels, LBO = the stride between the
            // two atoms (an image row of the halo tile, or the plane stride), K = 16 pixels = two 8-pixel swizzle atoms.
            constexpr uint32_t idesc_full = ptx::make_idesc_bf16(128, 128, 1, 1);
            constexpr uint32_t idesc_hi = ptx::make_idesc_bf16(128, 64, 1, 1);
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
                        const uint32_t in_row = in_base + rho * G::ROW_PAIR_BYTES + px_off;
                        const uint64_t g_desc = ptx::make_smem_desc_sw128(go_base + rho * G::ROW_PAIR_BYTES + px_off, G::PLANE_BYTES, 1024);
                        const uint64_t a01_hi = ptx::make_smem_desc_sw128(in_row, G::ROW_PAIR_BYTES, 1024);
                        const uint64_t a01_lo = ptx::make_smem_desc_sw128(in_row + G::PLANE_BYTES, G::ROW_PAIR_BYTES, 1024);
                        const uint64_t a2 = ptx::make_smem_desc_sw128(in_row + 2 * G::ROW_PAIR_BYTES, G::PLANE_BYTES, 1024);
                        ptx::umma_bf16(tmem_base, a01_hi, g_desc, idesc_full, accumulate);
                        ptx::umma_bf16(tmem_base, a01_lo, g_desc, idesc_hi, 1u);
                        ptx::umma_bf16(tmem_base + 128u, a2, g_desc, idesc_full, accumulate);
                        accumulate = 1;
                    }
                ptx::umma_commit(&bars->empty[st]);
                if (MC) ptx::umma_commit_multicast(&bars->gempty[st], (uint16_t)1);      // -> rank 0
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
            ptx::umma_commit(&bars->done);
        } else if (!(C == 64 && P3)) {
            // warp-uniform issue loop (see conv_tcp2.cu): all lanes run it, one elected lane issues
            auto issue_loop = [&](const bool el) {
                constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 192, 1, 1);
                constexpr uint32_t a_lbo = (C == 64) ? G::PLANE_BYTES : G::GO_CHUNK_BYTES;
                const uint32_t tb = __shfl_sync(__activemask(), tmem_base, 0);
                const uint32_t smem_u = ptx::smem_u32(smem);
                int st = 0; uint32_t ph = 0;
                uint32_t accumulate = 0;
#ifdef MSB_CONV_DEBUG
                long long dbg_wait = 0;
                const long long dbg_t0 = clock64();
#endif
                for (int tile = part; tile < num_tiles; tile += nparts) {
#ifdef MSB_CONV_DEBUG
                    { const long long t = clock64(); ptx::mbar_wait(&bars->full[st], ph); dbg_wait += clock64() - t; }
#else
                    ptx::mbar_wait(&bars->full[st], ph);
#endif
                    ptx::tc_fence_after();
                    const uint32_t go_base = smem_u + (uint32_t)(st * G::STAGE_BYTES);
                    const uint32_t in_base = go_base + G::GO_BYTES;
                    if (el) {
#pragma unroll
                        for (int rho = 0; rho < G::ROWS; ++rho)
#pragma unroll
                            for (int wg = 0; wg < WIMG / 16; ++wg) {
                                const uint32_t px_off = (uint32_t)wg * 16 * 128;
#pragma unroll
                                for (int pa = 0; pa < (C == 64 ? 1 : 2); ++pa) {      // C = 64: [g_hi ; g_lo] is one M = 128 operand
                                    const uint64_t adesc = ptx::make_smem_desc_sw128(
                                        go_base + rho * G::ROW_PAIR_BYTES + pa * G::PLANE_BYTES + px_off, a_lbo, 1024);
#pragma unroll
                                    for (int pb = 0; pb < 2; ++pb) {
                                        if (C != 64 && pa == 1 && pb == 1) continue;   // lo x lo (2^-18 relative) is dropped
                                        const uint64_t bdesc = ptx::make_smem_desc_sw128(
                                            in_base + rho * G::IN_ROW_PAIR_BYTES + pb * G::IN_PLANE_BYTES + px_off, G::B_LBO, 1024);
                                        ptx::umma_bf16(tb, adesc, bdesc, idesc, (rho | wg | pa | pb) ? 1u : accumulate);
                                    }
                                }
                            }
                        ptx::umma_commit(&bars->empty[st]);
                        if (MC) ptx::umma_commit_multicast(&bars->gempty[st], (uint16_t)1);      // -> rank 0
                    }
                    accumulate = 1;
                    if (++st == kStages) { st = 0; ph ^= 1; }
                }
                if (el) ptx::umma_commit(&bars->done);
#ifdef MSB_CONV_DEBUG
                if (el) {
                    dbg_loop_end = clock64();
                    atomicAdd(&g_wgrad_dbg[0], (unsigned long long)dbg_wait);
                    atomicAdd(&g_wgrad_dbg[1], (unsigned long long)(dbg_loop_end - dbg_t0));
                    atomicAdd(&g_wgrad_dbg[3], (unsigned long long)(dbg_t0 - dbg_entry));
                }
#endif
            };
            if (uniform_issue) {
                uint32_t e;
                asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(e));
                issue_loop(e != 0);
            } else if (lane == 0) {
                issue_loop(true);
            }
        }
    }
    // ===================== final pass: accumulator -> this CTA's partial =====================
    // Role-swapped C = 64 option: warps 4..7.  Everything else: ALL eight warps (the producer / issuer / allocator warps are
    // idle by now) -- warp w and warp w + 4 share TMEM quadrant w & 3 and take 96 of the 192 columns each, 32 columns (32
    // outstanding reads of the old partial) at a time: 3 dependent L2 round trips per thread instead of 12.
    __syncwarp();
    if (!(C == 64 && P3) || warp >= 4) {
        const int q = warp & 3;
        ptx::mbar_wait_backoff(&bars->done, 0, backoff_ns ? 8 * backoff_ns : 0);     // waits for the whole kernel
        ptx::tc_fence_after();
#ifdef MSB_CONV_DEBUG
        if (warp == 4 && lane == 0) dbg_epi_start = clock64();
#endif
        const int row = q * 32 + lane;                       // accumulator row
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (C == 64 && P3) {
            // final epilogue, C = 64: lane = (r or plane, c_in), columns = (g plane, c_out); the partials keep the layout
            // [slice][tap][c_in][c_out] -- a thread writes 16 consecutive c_out (64 B) per step.  Slice 2 part + 0 takes
            // everything except the in_lo rows of the third tap, which go to slice 2 part + 1 (zero elsewhere).
            const int ci = row & 63, half = row >> 6;
            float* const s0 = partial + (size_t)(part * 2) * 9 * C * C;
            float* const s1 = s0 + (size_t)9 * C * C;
#pragma unroll 1
            for (int cb = 0; cb < 64; cb += 16) {
                float a[16], b[16];
                // taps (r = half, s): in_hi x g_hi + in_lo x g_hi (columns cb ..) + in_hi x g_lo (columns 64 + cb ..)
                ptx::tmem_ld16(t_addr + cb, a);
                ptx::tmem_ld16(t_addr + 64 + cb, b);
                ptx::tmem_ld_wait();
                float* d01 = s0 + ((size_t)(half * 3 + s) * C + ci) * C + cb;
                float* z01 = s1 + ((size_t)(half * 3 + s) * C + ci) * C + cb;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = a[j] + b[j];
                    d01[j] = accumulate_partial ? d01[j] + v : v;
                    if (!accumulate_partial) z01[j] = 0.f;
                }
                // tap (r = 2, s): rows of the hi plane of `in` -> slice 0, of the lo plane -> slice 1
                ptx::tmem_ld16(t_addr + 128 + cb, a);
                ptx::tmem_ld16(t_addr + 192 + cb, b);
                ptx::tmem_ld_wait();
                float* d2 = (half ? s1 : s0) + ((size_t)(6 + s) * C + ci) * C + cb;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = a[j] + b[j];
                    d2[j] = accumulate_partial ? d2[j] + v : v;
                }
            }
        } else {
        // final epilogue: D[lane = gout channel (or hi/lo half x channel)][col = r*64 + ci] -> partial
        int slice, co;
        if (C == 64) { slice = part * 2 + (row >> 6); co = row & 63; }
        else { slice = part; co = row; }
        float* dst = partial + (size_t)slice * 9 * C * C;
        const int cb_begin = (warp >> 2) * 96;
#pragma unroll 1
        for (int cb = cb_begin; cb < cb_begin + 96; cb += 32) {
            float v[32];
            ptx::tmem_ld16(t_addr + cb, v);
            ptx::tmem_ld16(t_addr + cb + 16, v + 16);
            ptx::tmem_ld_wait();
            // the 32 columns lie inside one 64-column block = one tap (column block = the tap the N atom carries)
            const int tap = HT ? s * 3 + (cb >> 6) : (cb >> 6) * 3 + s;
            float* const q0 = dst + ((size_t)tap * C + ci_chunk * 64 + (cb & 63)) * C + co;      // element j: q0 + j * C
            // CTA-private slot: deterministic.  All old values are requested before the first store: left to itself the
            // compiler serialised load -> add -> store per element in one instantiation (it cannot rule out that the stores
            // alias the later loads), 192 dependent L2 round trips = 131k clocks instead of 11k.
            if (accumulate_partial) {
                float old[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) old[j] = __ldcg(q0 + (size_t)j * C);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += old[j];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) q0[(size_t)j * C] = v[j];
        }
        }
    }
#ifdef MSB_CONV_DEBUG
    if (warp == 4 && lane == 0) {      // [6] last MMA issued -> all MMAs complete, [7] the final TMEM -> global pass
        const long long now = clock64();
        // dbg_done is only defined inside the branch above: recompute the pieces from shared state
        atomicAdd(&g_wgrad_dbg[7], (unsigned long long)(now - dbg_epi_start));
        atomicAdd(&g_wgrad_dbg[6], (unsigned long long)(dbg_epi_start > dbg_loop_end ? dbg_epi_start - dbg_loop_end : 0));
    }
#endif
    ptx::tc_fence_before();
    __syncthreads();
#ifdef MSB_CONV_DEBUG
    if (threadIdx.x == 0) { atomicAdd(&g_wgrad_dbg[4], (unsigned long long)(clock64() - dbg_loop_end)); atomicAdd(&g_wgrad_dbg[5], 1ull); }
#endif
    if (MC) ptx::cluster_sync();         // nobody leaves while a peer may still multicast into / signal this CTA
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int C, int WIMG>
int nparts_impl(ConvShape s) {
    using G = WG<C, WIMG, true>;
    const int num_tiles = s.B * (s.H / G::ROWS);
    int np = num_sms() / G::GROUPS;
    if (np > num_tiles) np = num_tiles;
    if (np < 1) np = 1;
    return np;
}

template <int C, int WIMG, bool HT>
int launch_geom(const __nv_bfloat16* gout, const __nv_bfloat16* in, float* partial, int* nparts_out, int accumulate,
                ConvShape s, cudaStream_t st, bool mc, bool p3);

template <int C, int WIMG>
int launch_impl(const __nv_bfloat16* gout, const __nv_bfloat16* in, float* partial, int* nparts_out, int accumulate,
                ConvShape s, cudaStream_t st) {
    const bool mc = tune_get(TUNE_WGRAD_MULTICAST) != 0;
    const bool p3 = C != 64 || tune_get(TUNE_WGRAD64_PRODUCTS) == 3;
    if ((C == 64 && p3) || !tune_get(TUNE_WGRAD_HTAPS)) return launch_geom<C, WIMG, false>(gout, in, partial, nparts_out, accumulate, s, st, mc, p3);
    return launch_geom<C, WIMG, true>(gout, in, partial, nparts_out, accumulate, s, st, mc, p3);
}

template <int C, int WIMG, bool HT>
int launch_geom(const __nv_bfloat16* gout, const __nv_bfloat16* in, float* partial, int* nparts_out, int accumulate,
                ConvShape s, cudaStream_t st, bool mc, bool p3) {
    using G = WG<C, WIMG, HT>;
    CUtensorMap tm_go, tm_in;
    if (make_tmap_split5d(&tm_go, gout, s.B, s.H, s.W, s.C, WIMG, G::ROWS)) return -1;
    if (make_tmap_split5d(&tm_in, in, s.B, s.H, s.W, s.C, G::IN_W, G::IN_ROWS)) return -1;
    const size_t smem = (size_t)G::STAGES * G::STAGE_BYTES + sizeof(WBarriers) + 1024;
    // (HT never meets the role-swapped C = 64 three-product kernel: launch_impl sends that option to HT = false)
    constexpr bool kSwapOk = !(HT && C == 64);
    auto kern = mc ? ((p3 && kSwapOk) ? wgrad3x3_tc_kernel<C, WIMG, true, kSwapOk, HT> : wgrad3x3_tc_kernel<C, WIMG, true, false, HT>)
                   : ((p3 && kSwapOk) ? wgrad3x3_tc_kernel<C, WIMG, false, kSwapOk, HT> : wgrad3x3_tc_kernel<C, WIMG, false, false, HT>);
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
                                (uint32_t)tune_get(TUNE_WAIT_BACKOFF), tune_get(TUNE_UNIFORM_ISSUE) & 2);
    } else {
        le = launch_maybe_pdl(kern, np * G::GROUPS, kThreads, smem, st, tm_go, tm_in, partial, num_tiles, tiles_per_img, np,
                              accumulate, (uint32_t)tune_get(TUNE_WAIT_BACKOFF), tune_get(TUNE_UNIFORM_ISSUE) & 2);
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
