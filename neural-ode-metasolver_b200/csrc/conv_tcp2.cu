// tcgen05 engine, pixel-major form on a CTA PAIR (tcgen05.mma.cta_group::2): the 3x3 convolution of conv_tcp.cu with
// M = 256 -- two 128-pixel tiles, one per CTA of a 2-CTA cluster -- so that the weight operand is SHARED by the pair.
//
//   D[pixel][c_out] = sum_{tap, c_in} X[pixel + tap][c_in] * W[c_out][tap][c_in]
//
//   A (M = 256)  = activations: each CTA supplies ITS 128 pixels (one bf16 plane, K-major) from its own shared memory
//   B (N rows)   = weights: each CTA holds HALF of the N rows -- rank r: [W_hi of channel half r ; W_lo of the other half]
//                  (C rows).  The N = 2C product X_hi * [..] reads all of them; the N = C product X_lo * W_hi reads the
//                  first C/2 rows of each CTA (W_hi of both halves) from the SAME region (round 2; round 1 kept a
//                  separate copy: 3/2 of the bytes, 3 ring stages instead of 5).
//   D            = 128 TMEM lanes (own pixels) x 2C fp32 columns in each CTA:
//                  [hi half 0 | lo half 1 + (X_lo W_hi) half 1 | hi half 1 | lo half 0]; the epilogue adds a channel's two
//                  columns (resident variant: the round-1 layout [hi | lo])
//
// Why: with both operands in shared memory a single-CTA MMA is limited by the shared-memory operand feed (~72 B/clk):
// per k-step the single-CTA form reads 128 + 2C and 128 + C operand rows, the pair form 128 + C and 128 + C/2 -- for
// C = 64: 448 -> 352 rows (the N = 128 / N = 64 MMAs were the slowest point of the engine, profiles/conv_forms_r1.txt);
// and every weight tile crosses L2 -> SM once per PAIR instead of once per CTA.  Three hi/lo products as in conv_tcp.cu.
//
// Protocol (rank 0 = leader issues every MMA for the pair):
//   full barriers (x_full, w_full) live in the LEADER; both CTAs' TMA loads complete_tx on them
//       (cp.async.bulk.tensor...cta_group::2 with the leader's barrier address), the leader's producer arms them with
//       the bytes of both CTAs
//   empty barriers (x_empty, w_empty) and tmem_full are local to each CTA and signalled by the leader's
//       tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11)
//   tmem_empty lives in the leader; the epilogue warps of BOTH CTAs arrive on it (remote mbarrier.arrive)
// Epilogue: the 16-warp quad-transposed form of conv_tcp.cu (whole 128-byte lines per quad of lanes).
#include <cuda.h>

#include "msb_internal.h"
#include "msb_ptx.cuh"

#ifdef MSB_CONV_DEBUG
// instrumented build only: clocks the MMA-issuing warp spends in each of its waits, summed over the leader CTAs
// [0] weights full  [1] activations full  [2] accumulator empty  [3] whole issue loop  [4] leaders counted
static __device__ unsigned long long g_tcp2_wait[8];
extern "C" int msb_debug_tcp2_read(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_tcp2_wait, sizeof(g_tcp2_wait)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long zero[8] = {0};
        if (cudaMemcpyToSymbol(g_tcp2_wait, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#define TCP2_WAIT(bar, parity, slot) do { const long long _t = clock64(); ptx::mbar_wait(bar, parity); wait_clk[slot] += clock64() - _t; } while (0)
#else
#define TCP2_WAIT(bar, parity, slot) ptx::mbar_wait(bar, parity)
#endif

namespace msb {

int make_tmap_rows64(CUtensorMap* m, const void* base, size_t rows, int box_rows);
int make_tmap_split_plane(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h);
size_t tcp_packed_weight_bytes(int C);

namespace {

constexpr int kEpiWarp0 = 4;
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 230400;
constexpr int kStageBytesPerWarp = 4096;

// EW = epilogue warps: 16 = weight ring (any C); 8 = RESIDENT WEIGHTS (C = 64 only).  With N = 128 / 64 MMAs a weight
// stage is ~0.3 us of MMA work, shorter than the cross-CTA round trip its ring needs (commit -> multicast -> peer
// producer -> TMA -> remote complete_tx; profiles/conv_pair64_ab_r1.txt).  A CTA of a pair holds only HALF of every
// weight tile, so for C = 64 all nine taps (9 x 12 KB) fit next to the two activation stages: they are loaded once per
// kernel and the weight ring, its barriers and ~600 MB of L2 -> SM weight traffic per launch disappear.  The price is
// the shared memory of the epilogue stage: 8 warps with 2 KB each, transposing 16 pixels at a time.  Measured: 185 us per
// launch vs 197 us for the ring pair and 171 us for the single-CTA kernel (profiles/conv_pair64_resident_ab_r1.txt), so
// it is an option (tcp_epi_warps = 8 with tc_pair = 1), not the default.
// HS (ring form, 16 warps): HALF-size epilogue stage -- a warp transposes its 32 pixels in two passes of 16 (re-reading the
// accumulator from tensor memory for the second pass) -- and the 32 KB it frees buy a THIRD activation stage: the MMA warp's
// activation waits were what the 5-stage weight ring left over (scripts/diag_tcp2_waits.py).
// HALO (16-pixel-wide images whose height is a multiple of 16): ONE staged tile per c_in chunk serves all nine taps.
// tcgen05.mma takes a shared-memory operand whose start lies anywhere inside a 128-byte-swizzle atom and whose 8-row groups
// follow each other at any stride (the swizzle is a function of the absolute address bits; measured,
// scripts/probes/mma_rowshift_probe.cu), so a horizontal tap is the same tile read one pixel (128 B) further and needs no
// copy of its own.  For the 8-row groups of the M operand to sit at ONE stride, a CTA's 128 pixels are 16 image rows x
// 8 pixels (the left / right half of a 16 x 16 band; a pair = the band): group g = the 8 pixels of image row g, stride =
// one staged image row of 10 pixels (8 + a halo pixel each side).  Activation L2 -> SM traffic per tile: 2 planes x
// 18 x 10 pixels per chunk instead of 3 x (2 x 10 x 16): -62 %; a stage lasts 72 MMAs instead of 24.
constexpr int align1k(int x) { return (x + 1023) / 1024 * 1024; }
template <int C, int WIMG, int EW, bool HS = false, bool HALO = false> struct Geom2 {
    static constexpr bool RES = (EW == 8);
    static_assert(!(RES && (HS || HALO)), "half stage / halo tile: ring form only");
    static_assert(!RES || C == 64, "resident weights: C = 64 only");
    static_assert(!HALO || WIMG == 16, "halo tile: 16-pixel-wide images");
    static constexpr int ROWS = HALO ? 16 : 128 / WIMG;               // image rows of a CTA's tile
    static constexpr int TILE_W = HALO ? 8 : WIMG;                    // pixels per image row of a CTA's tile
    static constexpr int BOX_W = HALO ? TILE_W + 2 : WIMG;            // staged pixels per image row
    static constexpr int ROW_BYTES = BOX_W * 128;                     // one staged image row of one plane
    static constexpr int PLANE_BYTES = HALO ? align1k((ROWS + 2) * ROW_BYTES) : (ROWS + 2) * ROW_BYTES;
    static constexpr int X_STAGE_BYTES = 2 * PLANE_BYTES;
    static constexpr int A_SBO = HALO ? ROW_BYTES : 1024;             // stride between the 8-pixel groups of the M operand
    static constexpr int WA_BYTES = C * 128;                          // this CTA's half of [W_hi ; W_lo]
    static constexpr int WB_BYTES = (C / 2) * 128;                    // this CTA's half of W_hi
    // ring form: ONE region of C rows per CTA serves both MMAs of a k-step (see the weight producer); the resident
    // variant keeps the round-1 layout (region A + a separate half of W_hi for the lo-plane MMA)
    static constexpr int W_STAGE_BYTES = RES ? WA_BYTES + WB_BYTES : WA_BYTES;
    static constexpr int X_STAGES = HALO ? 2 : (HS ? 3 : 2);
    static constexpr int STAGE_BYTES = (RES || HS) ? EW * (kStageBytesPerWarp / 2) : EW * kStageBytesPerWarp;
    static constexpr int THREADS = (kEpiWarp0 + EW) * 32;
    static constexpr int RING = kSmemBudget - STAGE_BYTES - X_STAGES * X_STAGE_BYTES;
#ifdef MSB_TCP2_WSTAGES_CAP
    static constexpr int W_STAGES = RES ? 9 * (C / 64) : (RING / W_STAGE_BYTES > MSB_TCP2_WSTAGES_CAP ? MSB_TCP2_WSTAGES_CAP : RING / W_STAGE_BYTES);
#else
    static constexpr int W_STAGES = RES ? 9 * (C / 64) : (RING / W_STAGE_BYTES > kMaxStages ? kMaxStages : RING / W_STAGE_BYTES);
#endif
    static_assert(!RES || 9 * W_STAGE_BYTES <= RING, "resident weights do not fit");
    static constexpr int ACC_COLS = 2 * C;
    static constexpr int ACC_BUFS = (C == 64) ? 4 : 2;
    static_assert(ACC_COLS * ACC_BUFS <= 512, "TMEM");
    static_assert(W_STAGES >= 2, "weight ring");
};

struct __align__(8) Barriers2 {
    uint64_t w_full[kMaxStages], w_empty[kMaxStages];
    uint64_t x_full[kMaxStages], x_empty[kMaxStages];
    uint64_t tmem_full[4], tmem_empty[4];
    uint32_t tmem_base;
};

// tiles_per_img: 128-pixel tiles per image (HALO: 16-row bands = pairs per image)
template <int C, int WIMG, int ACT, int EW, bool HS, bool HALO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kEpiWarp0 + EW) * 32, 1)
conv3x3_tcp2_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                    const EpiParams epi, const int H, const int num_pairs, const int tiles_per_img,
                    const uint32_t backoff_ns, const int role_shift, const int uniform_issue) {
    using G = Geom2<C, WIMG, EW, HS, HALO>;
    constexpr int CHUNKS = C / 64;
    constexpr int kWStages = G::W_STAGES, kXStages = G::X_STAGES, kAccBufs = G::ACC_BUFS;
    constexpr int NG = C / 32, WPT = 4 * NG, TG = EW / WPT;            // epilogue: channel groups, warps per tile, tile groups
    static_assert(TG >= 1 && EW % WPT == 0, "the epilogue warps must split evenly over the channel groups");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_x = smem;
    uint8_t* smem_w = smem + kXStages * G::X_STAGE_BYTES;
    uint8_t* smem_stage = smem_w + kWStages * G::W_STAGE_BYTES;
    Barriers2* bars = reinterpret_cast<Barriers2*>(smem_stage + G::STAGE_BYTES);

    // role_shift = kEpiWarp0 (option mma_warp_high, default) rotates the roles so that the epilogue warps are the LOW
    // physical warps and TMA / MMA / alloc the last four: the schedulers favour the highest warp ids of a sub-partition,
    // and a late MMA issue is a tensor-pipe bubble while a late epilogue instruction is not.  (physical + 4) keeps
    // warp & 3, the TMEM lane quadrant a warp may read.
#ifdef MSB_CONV_DEBUG
    __shared__ unsigned long long dbg_epi_end, dbg_loop[2], dbg_t_entry;
    if (threadIdx.x == 0) { dbg_epi_end = 0; dbg_t_entry = (unsigned long long)clock64(); }
#endif
    const int warp = (int)(((threadIdx.x >> 5) + role_shift) % (kEpiWarp0 + EW));
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        ptx::prefetch_tmap(&tmap_w);
        for (int i = 0; i < (G::RES ? 1 : kWStages); ++i) { ptx::mbar_init(&bars->w_full[i], 1); ptx::mbar_init(&bars->w_empty[i], 1); }
        for (int i = 0; i < kXStages; ++i) { ptx::mbar_init(&bars->x_full[i], 1); ptx::mbar_init(&bars->x_empty[i], 1); }
        for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(&bars->tmem_full[i], 1); ptx::mbar_init(&bars->tmem_empty[i], 2 * WPT); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc2(&bars->tmem_base, kTmemCols);
        ptx::tmem_relinquish2();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                      // the peer's barriers are initialised before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    ptx::pdl_launch_dependents();
    // inputs are read (TMA, epilogue loads) and outputs written only after the predecessor has finished.  The weight
    // producer touches nothing but the packed weights: when the caller vouches that those were written by a fully ordered
    // launch further back (EpiParams::weights_settled, the ODE-block loops) it fills its ring while the predecessor's
    // last CTAs still run.
    if (!(warp == 3 && epi.weights_settled != 0)) ptx::pdl_wait();

    if (warp == 0) {
        // ===================== activation producer (both CTAs; own pixels) =====================
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int pr = cluster_id; pr < num_pairs; pr += num_clusters) {
                if constexpr (HALO) {
                    // one box per plane and chunk: image rows h0 - 1 .. h0 + 16, pixels 8 rank - 1 .. 8 rank + 8 (zero fill outside)
                    const int n = pr / tiles_per_img;
                    const int h0 = (pr - n * tiles_per_img) * G::ROWS;
                    for (int chunk = 0; chunk < CHUNKS; ++chunk) {
                        ptx::mbar_wait(&bars->x_empty[st], ph ^ 1);
                        const uint32_t full = ptx::mapa(ptx::smem_u32(&bars->x_full[st]), 0);
                        if (leader) ptx::mbar_arrive_expect_tx(&bars->x_full[st], 2 * 2 * (G::ROWS + 2) * G::ROW_BYTES);
                        uint8_t* dst = smem_x + st * G::X_STAGE_BYTES;
                        ptx::tma_load_5d_2sm(dst, &tmap_act, full, chunk * 64, (int)rank * G::TILE_W - 1, 0, h0 - 1, n);
                        ptx::tma_load_5d_2sm(dst + G::PLANE_BYTES, &tmap_act, full, chunk * 64, (int)rank * G::TILE_W - 1, 1, h0 - 1, n);
                        if (++st == kXStages) { st = 0; ph ^= 1; }
                    }
                    continue;
                }
                const int tile = 2 * pr + (int)rank;
                const int n = tile / tiles_per_img;
                const int h0 = (tile - n * tiles_per_img) * G::ROWS;
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s) {
                        ptx::mbar_wait(&bars->x_empty[st], ph ^ 1);
                        const uint32_t full = ptx::mapa(ptx::smem_u32(&bars->x_full[st]), 0);
                        if (leader) ptx::mbar_arrive_expect_tx(&bars->x_full[st], 2 * G::X_STAGE_BYTES);
                        uint8_t* dst = smem_x + st * G::X_STAGE_BYTES;
                        ptx::tma_load_5d_2sm(dst, &tmap_act, full, chunk * 64, s - 1, 0, h0 - 1, n);
                        ptx::tma_load_5d_2sm(dst + G::PLANE_BYTES, &tmap_act, full, chunk * 64, s - 1, 1, h0 - 1, n);
                        if (++st == kXStages) { st = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer (both CTAs; own halves) =====================
        if (lane == 0 && G::RES) {
            // resident: every tap once, this CTA's halves, all bytes of both CTAs counted on the leader's w_full[0]
            const uint32_t full = ptx::mapa(ptx::smem_u32(&bars->w_full[0]), 0);
            if (leader) ptx::mbar_arrive_expect_tx(&bars->w_full[0], 2 * 9 * G::W_STAGE_BYTES);
            for (int tap = 0; tap < 9; ++tap) {
                uint8_t* dst = smem_w + tap * G::W_STAGE_BYTES;
                const int row0 = tap * 2 * C;
                ptx::tma_load_2d_2sm(dst, &tmap_w, full, 0, row0 + (int)rank * C);
                ptx::tma_load_2d_2sm(dst + G::WB_BYTES, &tmap_w, full, 0, row0 + (int)rank * C + C / 2);
                ptx::tma_load_2d_2sm(dst + G::WA_BYTES, &tmap_w, full, 0, row0 + (int)rank * (C / 2));
            }
        } else if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int pr = cluster_id; pr < num_pairs; pr += num_clusters) {
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s)
                        for (int r = 0; r < 3; ++r) {
                            const int wt = (r * 3 + s) * CHUNKS + chunk;
                            ptx::mbar_wait(&bars->w_empty[st], ph ^ 1);
                            const uint32_t full = ptx::mapa(ptx::smem_u32(&bars->w_full[st]), 0);
                            if (leader) ptx::mbar_arrive_expect_tx(&bars->w_full[st], 2 * G::W_STAGE_BYTES);
                            uint8_t* dst = smem_w + st * G::W_STAGE_BYTES;
                            const int row0 = wt * 2 * C;                    // packed tile: rows [W_hi (C) ; W_lo (C)]
                            // rank r holds [W_hi of its channel half r ; W_lo of the OTHER half] (two boxes of C/2 rows).
                            // N = 2C MMA: columns [hi half 0 | lo half 1 | hi half 1 | lo half 0].  The N = C MMA of the lo
                            // activation plane reads the FIRST C/2 rows of each CTA = W_hi of both halves, and its columns
                            // [half 0 | half 1] land on accumulator columns that belong to the same channels -- no separate
                            // copy of W_hi: 2/3 of the bytes per stage, 5 stages instead of 3 in the same shared memory.
                            ptx::tma_load_2d_2sm(dst, &tmap_w, full, 0, row0 + (int)rank * (C / 2));
                            ptx::tma_load_2d_2sm(dst + G::WB_BYTES, &tmap_w, full, 0, row0 + C + (1 - (int)rank) * (C / 2));
                            if (++st == kWStages) { st = 0; ph ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the LEADER, for the pair) =====================
        // uniform_issue (default): the whole warp runs the loop -- waits, ring bookkeeping and descriptor arithmetic are
        // warp-uniform and live in uniform registers -- and one elected lane issues the MMAs / commits.  The round-1 style
        // (`if (lane == 0)` around everything) makes every operand of every MMA go through a ~15-instruction R2UR election
        // loop, which at 64 - 128 clocks per MMA and a pipe that queues only ~4 of them is close to the issue limit.
        auto issue_loop = [&](const bool el) {
            constexpr uint32_t idesc_full = ptx::make_idesc_bf16(256, 2 * C, 0, 0);   // X_hi * [W_hi ; W_lo]
            constexpr uint32_t idesc_hi = ptx::make_idesc_bf16(256, C, 0, 0);         // X_lo * W_hi
            const uint32_t tb = __shfl_sync(__activemask(), tmem_base, 0);
            const uint32_t xs_u32 = ptx::smem_u32(smem_x), ws_u32 = ptx::smem_u32(smem_w);
            int wst = 0, xst = 0; uint32_t wph = 0, xph = 0;
            int acc = 0; uint32_t acc_ph = 0;
#ifdef MSB_CONV_DEBUG
            long long wait_clk[3] = {0, 0, 0};
            const long long loop_t0 = clock64();
#endif
            if (G::RES) { ptx::mbar_wait(&bars->w_full[0], 0); ptx::tc_fence_after(); }
            for (int pr = cluster_id; pr < num_pairs; pr += num_clusters) {
                TCP2_WAIT(&bars->tmem_empty[acc], acc_ph ^ 1, 2);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tb + (uint32_t)(acc * G::ACC_COLS);
                uint32_t accumulate = 0;
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s) {
                        if (!HALO || s == 0) {                 // HALO: one staged tile per chunk serves the three horizontal taps
                            TCP2_WAIT(&bars->x_full[xst], xph, 1);
                            ptx::tc_fence_after();
                        }
                        const uint32_t x_base = xs_u32 + (uint32_t)(xst * G::X_STAGE_BYTES) + (HALO ? (uint32_t)s * 128u : 0u);
                        for (int r = 0; r < 3; ++r) {
                            if (G::RES) wst = r * 3 + s;
                            else { TCP2_WAIT(&bars->w_full[wst], wph, 0); ptx::tc_fence_after(); }
                            const uint32_t w_base = ws_u32 + (uint32_t)(wst * G::W_STAGE_BYTES);
                            if (el) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t wa = ptx::make_smem_desc_sw128(w_base + k * 32, 16, 1024);
                                    const uint64_t wb = ptx::make_smem_desc_sw128(w_base + (G::RES ? G::WA_BYTES : 0) + k * 32, 16, 1024);
                                    const uint64_t xhi = ptx::make_smem_desc_sw128(x_base + r * G::ROW_BYTES + k * 32, 16, G::A_SBO);
                                    const uint64_t xlo =
                                        ptx::make_smem_desc_sw128(x_base + G::PLANE_BYTES + r * G::ROW_BYTES + k * 32, 16, G::A_SBO);
                                    ptx::umma_bf16_2sm(d_tmem, xhi, wa, idesc_full, (k == 0) ? accumulate : 1u);
                                    ptx::umma_bf16_2sm(d_tmem, xlo, wb, idesc_hi, 1u);
                                }
                                if (!G::RES) ptx::umma_commit_2sm(&bars->w_empty[wst]);
                            }
                            accumulate = 1;
                            if (!G::RES) { if (++wst == kWStages) { wst = 0; wph ^= 1; } }
                        }
                        if (!HALO || s == 2) {
                            if (el) ptx::umma_commit_2sm(&bars->x_empty[xst]);
                            if (++xst == kXStages) { xst = 0; xph ^= 1; }
                        }
                    }
                if (el) ptx::umma_commit_2sm(&bars->tmem_full[acc]);
                if (++acc == kAccBufs) { acc = 0; acc_ph ^= 1; }
            }
#ifdef MSB_CONV_DEBUG
            if (el) {
                dbg_loop[0] = (unsigned long long)loop_t0;
                dbg_loop[1] = (unsigned long long)clock64();
                for (int i = 0; i < 3; ++i) atomicAdd(&g_tcp2_wait[i], (unsigned long long)wait_clk[i]);
                atomicAdd(&g_tcp2_wait[3], (unsigned long long)(clock64() - loop_t0));
                atomicAdd(&g_tcp2_wait[4], 1ull);
            }
#endif
        };
        if (leader) {
            if (uniform_issue) {
                uint32_t e;
                asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(e));
                issue_loop(e != 0);
            } else if (lane == 0) {
                issue_loop(true);
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue (both CTAs; see conv_tcp.cu, 16-warp form) =====================
        const int we = warp - kEpiWarp0;
        const int q = warp & 3;
        uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        asm volatile("" : "+r"(lane_addr));
        const size_t plane_stride = (size_t)WIMG * C;
        const int tg = we / WPT, wi = we % WPT;
        const int cb = (wi >> 2) * 32;
        // first global image row (n * H + h) of this CTA's tile of pair `pr`, its image, and pixel pp (0..127 = TMEM lane) of a tile
        auto row0_of = [&](int pr) {
            if constexpr (HALO) {
                const int n = pr / tiles_per_img;
                return (size_t)n * H + (size_t)(pr - n * tiles_per_img) * G::ROWS;
            } else {
                const int t = 2 * pr + (int)rank;
                const int n = t / tiles_per_img;
                return (size_t)n * H + (size_t)(t - n * tiles_per_img) * G::ROWS;
            }
        };
        auto img_of = [&](int pr) { return HALO ? pr / tiles_per_img : (2 * pr + (int)rank) / tiles_per_img; };
        auto pix_rc = [&](int pp, int& r2, int& w2) {
            r2 = pp / G::TILE_W;
            w2 = pp - r2 * G::TILE_W + (HALO ? (int)rank * G::TILE_W : 0);
        };
        if constexpr (G::RES) {
            // ---- resident-weight variant: 8 warps, 2 KB stage per warp, 16 pixels transposed at a time ----
            // lane pair (2m, 2m+1) owns pixel m of the current half: 64 B (16 channels) each, i.e. the pair covers the
            // pixel's whole 128-byte line; block b = 2*pass + qq is pixel 16*pass + (lane >> 1), channels 16*(lane & 1) + 8*qq
            const int cb = ((we % WPT) >> 2) * 32;
            const uint32_t stage = ptx::smem_u32(smem_stage + we * (kStageBytesPerWarp / 2));
            const int rl2 = lane >> 1, half = lane & 1;
            auto owned = [&](size_t row0, int b, size_t& idx, size_t& sidx) {
                const int pp = q * 32 + 16 * (b >> 1) + rl2;
                const int r2 = pp / WIMG, w2 = pp - r2 * WIMG;
                const int ch = cb + 16 * half + 8 * (b & 1);
                idx = ((row0 + r2) * WIMG + w2) * C + ch;
                sidx = ((row0 + r2) * 2) * plane_stride + (size_t)w2 * C + ch;
            };
            int it = 0;
            int pr = cluster_id;
            EpiVec8 ops;
            if (pr < num_pairs) { size_t i0, s0; owned(row0_of(pr), 0, i0, s0); epi_prefetch_vec8(epi, i0, ops); }
            for (; pr < num_pairs; pr += num_clusters, ++it) {
                const int acc = it % kAccBufs;
                const uint32_t ph = (uint32_t)(it / kAccBufs) & 1u;
                const EpiCoef coef = epi_coef(epi, img_of(pr));
                const size_t row0 = row0_of(pr);
                ptx::mbar_wait_backoff(&bars->tmem_full[acc], ph, backoff_ns);
                ptx::tc_fence_after();
                const uint32_t t_acc = tmem_base + (uint32_t)(acc * G::ACC_COLS) + lane_addr + (uint32_t)cb;
                float vv[4][8];
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float a[8], b[8];
                    ptx::tmem_ld<8>(t_acc + c8 * 8, a);
                    ptx::tmem_ld<8>(t_acc + C + c8 * 8, b);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) vv[c8][j] = a[j] + b[j];
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tmem_empty[acc]), 0));
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    __syncwarp();                                  // the previous half has been read by every lane
                    if ((lane >> 4) == pass) {
                        const int rl = lane & 15;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const uint32_t addr = stage + rl * 128 + ((u ^ (rl & 7)) << 4);
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(vv[u >> 1][4 * (u & 1)]),
                                         "f"(vv[u >> 1][4 * (u & 1) + 1]), "f"(vv[u >> 1][4 * (u & 1) + 2]),
                                         "f"(vv[u >> 1][4 * (u & 1) + 3]) : "memory");
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq) {
                        const int b = 2 * pass + qq;
                        float v[8];
                        {
                            const uint32_t base = stage + rl2 * 128;
                            const int u0 = 4 * half + 2 * qq;
                            const uint32_t a0 = base + ((u0 ^ (rl2 & 7)) << 4), a1 = base + (((u0 + 1) ^ (rl2 & 7)) << 4);
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a0));
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a1));
                        }
                        size_t idx, sidx;
                        owned(row0, b, idx, sidx);
                        epi_finish_v8<ACT>(epi, coef, v, ops, idx, sidx, plane_stride);
                        if (b < 3) {
                            owned(row0, b + 1, idx, sidx);
                            epi_prefetch_vec8(epi, idx, ops);
                        } else {
                            const int p2 = pr + num_clusters;
                            if (p2 < num_pairs) { owned(row0_of(p2), 0, idx, sidx); epi_prefetch_vec8(epi, idx, ops); }
                        }
                    }
                }
            }
        } else if constexpr (HS) {
        // ---- half-size stage: 16 pixels x 32 channels (2 KB) per pass, two passes per tile.  Pass p stages the pixels of
        //      lanes 16p..16p+15 (every lane re-reads its accumulator row, the other half discards it); the quads of lanes
        //      then read whole 128-byte lines as in the full-size form.  The accumulator is handed back after the second
        //      pass's read.  Step s = 2 * pass + jj owns pixel q*32 + 16*pass + 2*quad + jj. ----
        const uint32_t stage = ptx::smem_u32(smem_stage + we * (kStageBytesPerWarp / 2));
        const int k4 = lane & 3, quad = lane >> 2;
        auto owned = [&](size_t row0, int sidx_step, size_t& idx, size_t& sidx) {
            const int pp = q * 32 + 16 * (sidx_step >> 1) + quad * 2 + (sidx_step & 1);
            int r2, w2;
            pix_rc(pp, r2, w2);
            idx = ((row0 + r2) * WIMG + w2) * C + cb + 8 * k4;
            sidx = ((row0 + r2) * 2) * plane_stride + (size_t)w2 * C + cb + 8 * k4;
        };
        int it = tg;
        int pr = cluster_id + tg * num_clusters;
        EpiVec8 ops;
        if (pr < num_pairs) { size_t i0, s0; owned(row0_of(pr), 0, i0, s0); epi_prefetch_vec8(epi, i0, ops); }
        for (; pr < num_pairs; pr += TG * num_clusters, it += TG) {
            const int acc = it % kAccBufs;
            const uint32_t ph = (uint32_t)(it / kAccBufs) & 1u;
            const EpiCoef coef = epi_coef(epi, img_of(pr));
            const size_t row0 = row0_of(pr);
            ptx::mbar_wait_backoff(&bars->tmem_full[acc], ph, backoff_ns);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)(acc * G::ACC_COLS) + lane_addr + (uint32_t)cb;
            const uint32_t other = cb < C / 2 ? 3 * C / 2 : C / 2;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                __syncwarp();                                  // the previous half has been read by every lane
                const int rl = lane & 15;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float a[8], b[8];
                    ptx::tmem_ld<8>(t_acc + c8 * 8, a);
                    ptx::tmem_ld<8>(t_acc + other + c8 * 8, b);
                    ptx::tmem_ld_wait();
                    if ((lane >> 4) == pass) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const uint32_t addr = stage + rl * 128 + (((c8 * 2 + u) ^ (rl & 7)) << 4);
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a[4 * u] + b[4 * u]),
                                         "f"(a[4 * u + 1] + b[4 * u + 1]), "f"(a[4 * u + 2] + b[4 * u + 2]),
                                         "f"(a[4 * u + 3] + b[4 * u + 3]) : "memory");
                        }
                    }
                }
                if (pass == 1) ptx::tc_fence_before();
                __syncwarp();
                if (pass == 1 && lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tmem_empty[acc]), 0));
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int step = 2 * pass + jj;
                    const int row = quad * 2 + jj;
                    float v[8];
                    {
                        const uint32_t base = stage + row * 128;
                        const uint32_t a0 = base + (((2 * k4) ^ (row & 7)) << 4), a1 = base + (((2 * k4 + 1) ^ (row & 7)) << 4);
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a0));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a1));
                    }
                    size_t idx, sidx;
                    owned(row0, step, idx, sidx);
                    epi_finish_v8<ACT>(epi, coef, v, ops, idx, sidx, plane_stride);
                    if (step < 3) {
                        owned(row0, step + 1, idx, sidx);
                        epi_prefetch_vec8(epi, idx, ops);
                    } else {
                        const int p2 = pr + TG * num_clusters;
                        if (p2 < num_pairs) { owned(row0_of(p2), 0, idx, sidx); epi_prefetch_vec8(epi, idx, ops); }
                    }
                }
            }
        }
        } else {
        const uint32_t stage = ptx::smem_u32(smem_stage + we * kStageBytesPerWarp);
        const int k4 = lane & 3, quad = lane >> 2;
        auto owned = [&](size_t row0, int j, size_t& idx, size_t& sidx) {
            const int pp = q * 32 + quad * 4 + j;
            int r2, w2;
            pix_rc(pp, r2, w2);
            idx = ((row0 + r2) * WIMG + w2) * C + cb + 8 * k4;
            sidx = ((row0 + r2) * 2) * plane_stride + (size_t)w2 * C + cb + 8 * k4;
        };
        int it = tg;
        int pr = cluster_id + tg * num_clusters;
        EpiVec8 ops;
        if (pr < num_pairs) { size_t i0, s0; owned(row0_of(pr), 0, i0, s0); epi_prefetch_vec8(epi, i0, ops); }
        for (; pr < num_pairs; pr += TG * num_clusters, it += TG) {
            const int acc = it % kAccBufs;
            const uint32_t ph = (uint32_t)(it / kAccBufs) & 1u;
            const EpiCoef coef = epi_coef(epi, img_of(pr));
            const size_t row0 = row0_of(pr);
            ptx::mbar_wait_backoff(&bars->tmem_full[acc], ph, backoff_ns);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)(acc * G::ACC_COLS) + lane_addr + (uint32_t)cb;
            // the second column set of channel block cb (accumulator columns [hi half 0 | lo half 1 | hi half 1 | lo half 0])
            const uint32_t other = cb < C / 2 ? 3 * C / 2 : C / 2;
            __syncwarp();
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float a[8], b[8];
                ptx::tmem_ld<8>(t_acc + c8 * 8, a);
                ptx::tmem_ld<8>(t_acc + other + c8 * 8, b);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t addr = stage + lane * 128 + (((c8 * 2 + u) ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a[4 * u] + b[4 * u]),
                                 "f"(a[4 * u + 1] + b[4 * u + 1]), "f"(a[4 * u + 2] + b[4 * u + 2]),
                                 "f"(a[4 * u + 3] + b[4 * u + 3]) : "memory");
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->tmem_empty[acc]), 0));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = quad * 4 + j;
                float v[8];
                {
                    const uint32_t base = stage + row * 128;
                    const uint32_t a0 = base + (((2 * k4) ^ (row & 7)) << 4), a1 = base + (((2 * k4 + 1) ^ (row & 7)) << 4);
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a0));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a1));
                }
                size_t idx, sidx, nidx = 0, nsidx;
                owned(row0, j, idx, sidx);
                // the next group's operand loads are issued from inside the finish, right after this group's operands are consumed
                bool has_next = true;
                if (j < 3) owned(row0, j + 1, nidx, nsidx);
                else {
                    const int p2 = pr + TG * num_clusters;
                    has_next = p2 < num_pairs;
                    if (has_next) owned(row0_of(p2), 0, nidx, nsidx);
                }
                epi_finish_v8<ACT>(epi, coef, v, ops, idx, sidx, plane_stride,
                                   [&]() { if (has_next) epi_prefetch_vec8(epi, nidx, ops); });
            }
        }
        }
    }
#ifdef MSB_CONV_DEBUG
    if (warp >= kEpiWarp0 && lane == 0) atomicMax(&dbg_epi_end, (unsigned long long)clock64());
#endif
    // nobody leaves while the peer may still signal its barriers or the leader's MMAs write its TMEM
    ptx::tc_fence_before();
    __syncthreads();
#ifdef MSB_CONV_DEBUG
    if (leader && threadIdx.x == 0) {          // [5] entry -> MMA loop start, [6] MMA loop end -> last epilogue warp done, [7] entry -> here
        atomicAdd(&g_tcp2_wait[5], dbg_loop[0] - dbg_t_entry);
        atomicAdd(&g_tcp2_wait[6], dbg_epi_end > dbg_loop[1] ? dbg_epi_end - dbg_loop[1] : 0ull);
        atomicAdd(&g_tcp2_wait[7], (unsigned long long)clock64() - dbg_t_entry);
    }
#endif
    ptx::cluster_sync();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, kTmemCols);
    }
}

template <int C, int WIMG, int ACT, int EW, bool HS = false, bool HALO = false>
int launch_act_ew(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
                  cudaStream_t st) {
    using G = Geom2<C, WIMG, EW, HS, HALO>;
    constexpr int kThreads = G::THREADS;
    CUtensorMap tm_act, tm_w;
    if (make_tmap_split_plane(&tm_act, split_in, s.B, s.H, s.W, s.C, G::BOX_W, G::ROWS + 2)) return -1;
    if (make_tmap_rows64(&tm_w, w_tiles, tcp_packed_weight_bytes(C) / 128, C / 2)) return -1;
    constexpr size_t smem = (size_t)G::X_STAGES * G::X_STAGE_BYTES + (size_t)G::W_STAGES * G::W_STAGE_BYTES + G::STAGE_BYTES +
                            sizeof(Barriers2) + 1024;
    static_assert(smem <= 232448, "shared memory per CTA");
    auto kern = conv3x3_tcp2_kernel<C, WIMG, ACT, EW, HS, HALO>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(conv3x3_tcp2)"))
        return -1;
    const int tiles_per_img = s.H / G::ROWS;                                        // HALO: bands = pairs per image
    const int num_pairs = HALO ? s.B * tiles_per_img : s.B * tiles_per_img / 2;
    const int clusters = std::min(num_pairs, num_sms() / 2);
    const cudaError_t le = launch_maybe_pdl(kern, 2 * clusters, kThreads, smem, st, tm_act, tm_w, epi, s.H, num_pairs,
                                            tiles_per_img, (uint32_t)tune_get(TUNE_WAIT_BACKOFF), tune_get(TUNE_MMA_WARP_HIGH) ? kEpiWarp0 : 0,
                                            tune_get(TUNE_UNIFORM_ISSUE) & 1);
    count_launch();
    return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "conv3x3_tcp2 launch");
}

template <int C, int WIMG, int ACT>
int launch_act(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
               cudaStream_t st) {
    if constexpr (C == 64) {
        if (tune_get(TUNE_TCP_EPI_WARPS) == 8) return launch_act_ew<C, WIMG, ACT, 8>(split_in, w_tiles, epi, s, st);
    }
    if constexpr (C == 128 && WIMG == 16) {
        if (tune_get(TUNE_TCP2_HALO) && s.H % 16 == 0) {
            if (tune_get(TUNE_TCP2_HALF_STAGE)) return launch_act_ew<C, WIMG, ACT, 16, true, true>(split_in, w_tiles, epi, s, st);
            return launch_act_ew<C, WIMG, ACT, 16, false, true>(split_in, w_tiles, epi, s, st);
        }
    }
    if constexpr (C == 128) {
        if (tune_get(TUNE_TCP2_HALF_STAGE)) return launch_act_ew<C, WIMG, ACT, 16, true>(split_in, w_tiles, epi, s, st);
    }
    return launch_act_ew<C, WIMG, ACT, 16>(split_in, w_tiles, epi, s, st);
}

template <int C, int WIMG>
int launch_impl(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
                cudaStream_t st) {
    const int act = (epi.out_split || epi.dact_out) ? epi.act : ACT_NONE;
    if (act == ACT_GELU) return launch_act<C, WIMG, ACT_GELU>(split_in, w_tiles, epi, s, st);
    if (act == ACT_RELU) return launch_act<C, WIMG, ACT_RELU>(split_in, w_tiles, epi, s, st);
    return launch_act<C, WIMG, ACT_NONE>(split_in, w_tiles, epi, s, st);
}

}  // namespace

// the pair form needs an even number of 128-pixel tiles (pairs never straddle... they may: a pair is two consecutive
// tiles of the batch-major tile order, each CTA addresses its own tile independently)
bool tcp2_shape_supported(int B, int C, int H, int W) {
    if (!tc_shape_supported(C, H, W)) return false;
    const int rows = 128 / W;
    return ((B * (H / rows)) % 2) == 0;
}

int launch_conv3x3_tcp2(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
                        cudaStream_t st) {
    if (!tcp2_shape_supported(s.B, s.C, s.H, s.W)) {
        set_error("tcgen05 pair conv: unsupported shape B=%d C=%d H=%d W=%d", s.B, s.C, s.H, s.W);
        return -1;
    }
    if (epi.chan_bias || epi.pix_bias) { set_error("tcgen05 conv: bias terms are SIMT-engine only"); return -1; }
    if (s.C == 64 && s.W == 32) return launch_impl<64, 32>(split_in, w_tiles, epi, s, st);
    if (s.C == 64 && s.W == 16) return launch_impl<64, 16>(split_in, w_tiles, epi, s, st);
    if (s.C == 128 && s.W == 32) return launch_impl<128, 32>(split_in, w_tiles, epi, s, st);
    return launch_impl<128, 16>(split_in, w_tiles, epi, s, st);
}

}  // namespace msb
