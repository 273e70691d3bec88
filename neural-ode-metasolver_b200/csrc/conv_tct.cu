// tcgen05 engine, TMEM-resident-weight form ("tct") for C = 64: 3x3 (stride 1, pad 1) convolution as an implicit GEMM
// whose A operand (the weights) is read from TENSOR MEMORY and never leaves it for the lifetime of the kernel.
//
//   D^T[c_out][pixel] = sum_{tap, c_in} W[c_out][tap][c_in] * X[pixel + tap][c_in]
//
// Why: with both operands in shared memory the C = 64 MMAs (N = 128 / 64 in the pixel-major form, conv_tcp.cu) need
// ~150 B/clk of operand reads at the tensor-pipe rate while the shared-memory feed sustains ~64-72 B/clk (measured, round
// 1: MMA-only 131 us against 49 us at the clock peak), and every CTA re-streamed the same nine weight tiles for every
// tile (604 MB L2 -> SM per launch for a 147 KB tensor).  Here
//   A (M = 128 rows) = [W_hi ; W_lo] for all 9 taps x 64 c_in = 128 lanes x 288 TMEM columns (2 bf16 per column),
//                      written ONCE per CTA with tcgen05.st; each MMA names its K = 16 slab as a TMEM address;
//   B (N = 64 rows)  = one image row of activations, [X_hi (32 px) ; X_lo (32 px)], K-major, 128B-swizzled, from a
//                      ring of image rows in shared memory (each input row is loaded once per horizontal tap and
//                      serves the three vertical taps of three output rows);
//   D                = 128 lanes x 64 fp32 columns per output image row, three buffers.
// so the only shared-memory operand traffic is B: 64 rows x 32 B per 32-clock MMA = 64 B/clk, and the L2 -> SM operand
// traffic drops from 1.2 GB to ~0.44 GB per launch (B = 512).  All four hi/lo products are formed (M = 128 is the
// minimum full-rate M, and 64 output channels x {hi, lo} fill it exactly).
//
// Epilogue: lanes are (channel, hi/lo) rows, so the accumulator is first reduced (hi + lo columns in registers, hi + lo
// rows with one shuffle -- the rows of one channel sit 16 lanes apart inside a TMEM quadrant) and transposed through an
// XOR-swizzled 8 KB shared-memory stage per warp group, then finished in the pixel-major vector form (a thread owns 8
// consecutive channels of one pixel: 256-bit global accesses, 4 pixels = 1 KB contiguous per warp instruction) with the
// same arithmetic as every other engine (epi_finish_vec8).  16 epilogue warps = 4 groups of 4 (one per TMEM lane
// quadrant); group g takes output rows g, g+4, ...
//
// Decomposition switches (option tct_debug, a kernel PARAMETER so the release build can be taken apart without slowing
// it down): 1 = epilogue skipped (accumulator hand-shake still cycles), 2 = no MMAs issued, 8 = no activation TMA.
//
// Warp roles: 0..15 = epilogue, 16 = activation TMA (+ optional L2 prefetch of the epilogue operands), 17 = TMEM alloc,
// 18 / 19 = MMA issue (alternate rows).  Persistent: grid = min(#bands, #SMs); work item = a band of BH image rows of one image.
#include <cuda.h>

#include "msb_internal.h"
#include "msb_ptx.cuh"

namespace msb {

int make_tmap_split5d(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h);

bool tct_shape_supported(int C, int H, int W) { return C == 64 && W == 32 && H % 4 == 0; }
size_t tct_packed_weight_bytes() { return (size_t)128 * 288 * 4; }

#ifdef MSB_CONV_DEBUG
// instrumented build: [0] = bit mask of waits that timed out (bit = wait id), [1..] progress counters
static __device__ unsigned g_tct_dbg[16];
int tct_debug_set(int flags) {
    unsigned zero[16] = {0};
    for (int i = 8; i < 12; ++i) zero[i] = 0xffffffffu;
    if (cudaMemcpyToSymbol(g_tct_dbg, zero, sizeof(zero)) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(g_conv_debug, &flags, sizeof(int)) == cudaSuccess ? 0 : -1;
}
extern "C" int msb_debug_tct_read(unsigned* out16) {
    return cudaMemcpyFromSymbol(out16, g_tct_dbg, sizeof(g_tct_dbg)) == cudaSuccess ? 0 : -1;
}
// kernel timeline, summed over CTAs (clocks): [0] entry -> past griddepcontrol.wait, [1] -> weights in TMEM, [2] -> MMA issue
// loops done, [3] -> last epilogue warp done, [4] -> exit, [5] CTAs counted
static __device__ unsigned long long g_tct_time[12];
// [6..9] clocks the two MMA-issuing warps spend waiting for: a free accumulator, input rows, their turn; and issuing
extern "C" int msb_debug_tct_time(unsigned long long* out12, int reset) {
    if (out12 && cudaMemcpyFromSymbol(out12, g_tct_time, sizeof(g_tct_time)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long zero[12] = {0};
        if (cudaMemcpyToSymbol(g_tct_time, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#define TCT_STAMP(i) do { if (lane == 0) atomicMax(&dbg_t[i], (unsigned long long)clock64()); } while (0)
#define TCT_TIMED(slot, stmt) do { const long long _t = clock64(); stmt; dbg_w[slot] += clock64() - _t; } while (0)
// bounded wait that records its id and gives up (the kernel then finishes with garbage instead of trapping)
__device__ __forceinline__ void tct_wait(uint64_t* bar, uint32_t parity, int id, int row) {
    uint32_t spins = 0;
    while (!ptx::mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 16)) {
            atomicOr(&g_tct_dbg[0], 1u << id);
            atomicMin(&g_tct_dbg[8 + id], (unsigned)row);        // first row whose wait `id` timed out
            atomicMax(&g_tct_dbg[12 + id], (unsigned)row);       // last one
            return;
        }
    }
}
#define TCT_WAIT(bar, parity, id) tct_wait(bar, parity, id, TCT_ROW)
#define TCT_MARK(slot) atomicAdd(&g_tct_dbg[slot], 1u)
#else
#define TCT_WAIT(bar, parity, id) ptx::mbar_wait(bar, parity)
#define TCT_MARK(slot) ((void)0)
#define TCT_STAMP(i) ((void)0)
#define TCT_TIMED(slot, stmt) do { stmt; } while (0)
#endif

namespace {

// Warp roles.  The schedulers favour the HIGHEST warp ids of a sub-partition, so the two MMA-issuing warps come last:
// an MMA that is issued late is a tensor-pipe bubble (the pipe queues only ~4 instructions), an epilogue instruction
// that is issued late is not.
constexpr int kEpiWarps = 16;                              // warps 0..15: warp & 3 = TMEM lane quadrant, warp >> 2 = group
constexpr int kProducerWarp = 16;
constexpr int kAllocWarp = 17;
constexpr int kMmaWarp0 = 18;                              // warps 18, 19: even / odd output rows
constexpr int kThreads = 20 * 32;                          // 640
constexpr int kRowBytes = 2 * 32 * 128;                    // one image row, both planes (N = 64 rows of 128 B)
constexpr int kSlotBytes = 3 * kRowBytes;                  // ... for the three horizontal taps
constexpr int kSlots = 8;                                  // image rows in the ring
constexpr int kStageBytes = 32 * 64 * 4;                   // one output row, fp32, per epilogue group
constexpr int kMaxAccBufs = 6;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kWCol0 = 192;                           // accumulators: columns 0 .. 191, weights: columns 192 .. 479
// P3 (three hi/lo products, option tct_products = 3): per k-step  D[:, 32 px] (+)= [W_hi ; W_lo] * X_hi   (M = 128, N = 32)
//                                                     D[hi rows, 32 px] += W_hi * X_lo          (M = 64,  N = 32)
// An M = 64 MMA reads its A rows from / adds its D rows into lanes 0-15 of each 32-lane TMEM quadrant (measured,
// scripts/probes/mma_m64_probe.cu) -- exactly where the interleaved row order keeps the W_hi rows, so both MMAs name
// the same A slab and the same accumulator.  Same tensor-pipe time as the four-product form (an M = 64 MMA is not
// faster than an M = 128 one) but a quarter fewer MACs, which is what counts under the board power cap; the
// accumulator of an output row shrinks to 32 columns (the two activation planes are summed by the tensor core),
// so six rows are in flight instead of three.  Measured (profiles/conv_forms_r2.txt): 148 vs 121 us per launch -- 72 MMAs
// of 16 clocks per row are bound by the issue rate of the pipe, not by power -- so it is an option, not the default.
// !P3 (four products, the default): one M = 128, N = 64 MMA per k-step on [X_hi ; X_lo], 64-column accumulators, three in
// flight.
template <bool P3> struct AccGeom {
    static constexpr int BUFS = P3 ? 6 : 3;
    static constexpr uint32_t COLS = P3 ? 32 : 64;
};
constexpr int kWChunks = 18;                               // 288 columns in chunks of 16

// An mbarrier wait names a phase by its PARITY, so a waiter must see every phase of a barrier (it may lag by at most
// one).  Output row i uses accumulator i % 3 and epilogue group i % 4: a group meets a given accumulator only every
// 12 rows, so "accumulator full" gets one barrier per (accumulator, group) combination = row index mod 12 (each
// always waited on by the same group, phase = i / 12); "accumulator empty" has one waiter (the MMA warp, every
// phase in order) and stays per accumulator.
constexpr int kFullBars = 12;
struct __align__(8) Barriers {
    uint64_t full[kSlots], empty[kSlots];
    uint64_t tmem_full[kFullBars], tmem_empty[kMaxAccBufs];
    uint32_t tmem_base;
};

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int ACT, bool LEAN, bool P3>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tct_kernel(const __grid_constant__ CUtensorMap tmap_act, const uint4* __restrict__ wpacked, const EpiParams epi,
                   const int H, const int num_items, const int bands_per_img, const int BH, const int l2pf,
                   const uint32_t backoff_ns, const int dbg) {
    constexpr int C = 64, W = 32;
    constexpr int kAccBufs = AccGeom<P3>::BUFS;
    constexpr uint32_t kAccCols = AccGeom<P3>::COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_rows = smem;                                       // kSlots x kSlotBytes
    uint8_t* smem_stage = smem + kSlots * kSlotBytes;                // 4 x kStageBytes
    Barriers* bars = reinterpret_cast<Barriers*>(smem_stage + 4 * kStageBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
#ifdef MSB_CONV_DEBUG
    __shared__ unsigned long long dbg_t[5];
    if (threadIdx.x == 0) { dbg_t[0] = (unsigned long long)clock64(); dbg_t[1] = dbg_t[2] = dbg_t[3] = dbg_t[4] = 0; }
#endif

    if (warp == kProducerWarp && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        for (int i = 0; i < kSlots; ++i) { ptx::mbar_init(&bars->full[i], 1); ptx::mbar_init(&bars->empty[i], 1); }
        for (int i = 0; i < kFullBars; ++i) ptx::mbar_init(&bars->tmem_full[i], 1);
        for (int i = 0; i < kMaxAccBufs; ++i) ptx::mbar_init(&bars->tmem_empty[i], 4);
        ptx::fence_barrier_init();
    }
    if (warp == kAllocWarp) {
        ptx::tmem_alloc(&bars->tmem_base, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    ptx::pdl_launch_dependents();       // the next kernel may set itself up while this one runs ...

    // ---- weights -> TMEM, once (epilogue warps): thread (quadrant wq, lane) owns TMEM lane 32 wq + lane; the 18 column
    // chunks of its row are spread over the four groups.  Packed layout: [chunk][lane 0..127][16 x u32] (coalesced).
    auto load_weights = [&]() {
        const int g = warp >> 2, wq = warp & 3;
        const int L = wq * 32 + lane;
        auto load_chunk = [&](int ch, uint32_t* r) {
            const uint4* src = wpacked + ((size_t)ch * 128 + L) * 4;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 t = __ldg(src + v);
                r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
            }
        };
        const uint32_t w_lane = tmem_base + ((uint32_t)(wq * 32) << 16) + kWCol0;
        {
            // all (up to five) chunks of this thread are requested before the first is stored: one L2 round trip
            // instead of two (nothing else is live in registers yet)
            uint32_t r0[16], r1[16], r2[16], r3[16], r4[16];
            load_chunk(g, r0); load_chunk(g + 4, r1); load_chunk(g + 8, r2); load_chunk(g + 12, r3);
            if (g + 16 < kWChunks) load_chunk(g + 16, r4);
            tmem_st16(w_lane + (uint32_t)g * 16u, r0);
            tmem_st16(w_lane + (uint32_t)(g + 4) * 16u, r1);
            tmem_st16(w_lane + (uint32_t)(g + 8) * 16u, r2);
            tmem_st16(w_lane + (uint32_t)(g + 12) * 16u, r3);
            if (g + 16 < kWChunks) tmem_st16(w_lane + (uint32_t)(g + 16) * 16u, r4);
        }
        tmem_st_wait();
        ptx::tc_fence_before();
    };
    // The packed weights are the one input that the predecessor kernel did not write when the caller says so
    // (EpiParams::weights_settled: they were packed by a plain, fully ordered launch at least two launches back -- the
    // ODE-block loops).  They are then fetched BEFORE griddepcontrol.wait, while the predecessor's last CTAs still run.
    const bool w_early = epi.weights_settled != 0;
    if (w_early && warp < kEpiWarps) load_weights();
    ptx::pdl_wait();                    // ... and this one touches its other inputs only after its predecessor has finished
    if (warp == 0) TCT_STAMP(1);

    const int my_items = blockIdx.x < num_items ? (num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int my_rows = my_items * BH;

    if (warp == kProducerWarp) {
        // ===================== activation producer =====================
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int n = item / bands_per_img;
                const int h0 = (item - n * bands_per_img) * BH;
                for (int rr = -1; rr <= BH; ++rr) {
#define TCT_ROW (h0 + rr)
                    TCT_WAIT(&bars->empty[slot], ph ^ 1, 0);
#undef TCT_ROW
                    if (dbg & 8) { ptx::mbar_arrive(&bars->full[slot]); }
                    else {
                    ptx::mbar_arrive_expect_tx(&bars->full[slot], kSlotBytes);
                    uint8_t* dst = smem_rows + slot * kSlotBytes;
#pragma unroll
                    for (int s = 0; s < 3; ++s)
                        ptx::tma_load_5d(dst + s * kRowBytes, &tmap_act, &bars->full[slot], 0, s - 1, 0, h0 + rr, n);
                    }
                    if (++slot == kSlots) { slot = 0; ph ^= 1; }
                    // the epilogue's fp32 operands of an output row (one contiguous 8 KB range each) are pulled into L2
                    // about a ring depth ahead of their use
                    if (l2pf && rr >= 0 && rr < BH) {
                        const size_t off = ((size_t)n * H + h0 + rr) * W * C;
#pragma unroll
                        for (int i = 0; i < kEpiLoadSlots; ++i) {
                            const float* p = epi_load_operand(epi, i);
                            if (p) ptx::l2_prefetch_bulk(p + off, W * C * 4);
                        }
                    }
                }
            }
        }
    } else if (warp >= kMmaWarp0) {
        // ===================== MMA issuers (two warps, one elected thread each) =====================
        // The tensor pipe queues only a few MMAs (measured, scripts/probes/mma_sync_probe.cu: 36 N = 64 MMAs are issued in
        // 1276 clocks against 1152 of execution, and a wait + commit between rows costs a ~220-clock bubble), so ONE
        // issuing warp loses the per-row synchronisation time outright.  Two warps alternate rows: while one is blocked
        // issuing row i, the other has already done the waits of row i + 1 and starts issuing the moment it is handed
        // the turn (named barriers 6 / 7: "my row is issued").  Rows are therefore issued -- and, the pipe being in
        // order, completed -- in row order, which is what lets the commit of row i release input row i - 1.
        // Each warp runs its loop warp-uniformly (descriptors in uniform registers), one elected lane issues.
        const int w = warp - kMmaWarp0;
        named_bar_sync(5, (kEpiWarps + 2) * 32);       // the weights are in TMEM (written by the epilogue warps below)
        ptx::tc_fence_after();
        TCT_STAMP(2);
        {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, P3 ? 32 : 64, 0, 0);
            constexpr uint32_t idesc_hi = ptx::make_idesc_bf16(64, 32, 0, 0);       // P3: the W_hi rows only
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t w_tmem = tb + kWCol0;
            const uint32_t rows_u32 = ptx::smem_u32(smem_rows);
            const bool leader = elect_one();
            uint32_t q0 = 0, confirmed = 0;          // ring index of the first row of the current item / rows seen full
            int j = w;                               // row of the item (BH >= 4 > w)
#ifdef MSB_CONV_DEBUG
            long long dbg_w[4] = {0, 0, 0, 0};
#endif
            for (int i = w; i < my_rows; i += 2) {
                const int acc = i % kAccBufs;
                const uint32_t acc_ph = (uint32_t)(i / kAccBufs) & 1u;
#define TCT_ROW i
                TCT_TIMED(0, TCT_WAIT(&bars->tmem_empty[acc], acc_ph ^ 1, 1));
                ptx::tc_fence_after();
                const uint32_t d_tmem = tb + (uint32_t)acc * kAccCols;
                uint32_t b_base[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const uint32_t qq = q0 + (uint32_t)(j + r);
                    const uint32_t sl = qq % kSlots;
                    if (qq >= confirmed) {
                        TCT_TIMED(1, TCT_WAIT(&bars->full[sl], (qq / kSlots) & 1u, 2));
                        ptx::tc_fence_after();
                        confirmed = qq + 1;
                    }
                    b_base[r] = rows_u32 + sl * kSlotBytes;
                }
#undef TCT_ROW
                if (i > 0) TCT_TIMED(2, named_bar_sync(w == 0 ? 7 : 6, 64));      // the other warp has issued row i - 1
#ifdef MSB_CONV_DEBUG
                const long long dbg_issue0 = clock64();
#endif
                if (leader) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_base[r] + s * kRowBytes + k * 32, 16, 1024);
                                const uint32_t a_tmem = w_tmem + (uint32_t)((r * 3 + s) * 32 + k * 8);
                                if (!(dbg & 2)) {
                                    umma_bf16_ts(d_tmem, a_tmem, bdesc, idesc, (r | s | k) ? 1u : 0u);
                                    if (P3) {       // the lo plane of the activations starts 32 rows (4 KB) further
                                        const uint64_t bdesc_lo =
                                            ptx::make_smem_desc_sw128(b_base[r] + s * kRowBytes + kRowBytes / 2 + k * 32, 16, 1024);
                                        umma_bf16_ts(d_tmem, a_tmem, bdesc_lo, idesc_hi, 1u);
                                    }
                                }
                            }
                        }
                    }
                }
                __syncwarp();
#ifdef MSB_CONV_DEBUG
                dbg_w[3] += clock64() - dbg_issue0;
#endif
                if (i + 1 < my_rows) named_bar_arrive(w == 0 ? 6 : 7, 64);
                if (leader) {
                    ptx::umma_commit(&bars->tmem_full[i % kFullBars]);
                    ptx::umma_commit(&bars->empty[(q0 + (uint32_t)j) % kSlots]);
                    if (j == BH - 1) {
                        ptx::umma_commit(&bars->empty[(q0 + (uint32_t)j + 1) % kSlots]);
                        ptx::umma_commit(&bars->empty[(q0 + (uint32_t)j + 2) % kSlots]);
                    }
                }
                __syncwarp();
                j += 2;
                if (j >= BH) { j -= BH; q0 += (uint32_t)(BH + 2); }
            }
#ifdef MSB_CONV_DEBUG
            if (leader) for (int q = 0; q < 4; ++q) atomicAdd(&g_tct_time[6 + q], (unsigned long long)dbg_w[q]);
#endif
        }
    } else if (warp < kEpiWarps) {
        // ===================== epilogue =====================
        const int g = warp >> 2;                                // group: output rows g, g + 4, ...
        const int wq = warp & 3;                              // TMEM lane quadrant of this warp
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;

        if (!w_early) load_weights();
        {
            named_bar_sync(5, (kEpiWarps + 2) * 32);          // with the two MMA warps
        }

        const uint32_t stage = ptx::smem_u32(smem_stage + g * kStageBytes);
        const bool up = lane >= 16;                           // lo rows: keep pixels 16..31
        const int c = 16 * wq + (lane & 15);
        // writer: word(px, c) = px * 64 + (c ^ ((px >> 4) << 4) ^ (((c >> 5) & 1) << 2))
        const uint32_t st_addr = stage + 4u * (uint32_t)((up ? 16 : 0) * 64 + (c ^ (up ? 16 : 0) ^ ((wq >> 1) << 2)));
        // reader: pixel rp (+ 16 per pass), channels 8 k8 .. 8 k8 + 7
        const int k8 = lane & 7;
        const int rp = wq * 4 + (lane >> 3);
        const size_t plane_stride = (size_t)W * C;

        // row i of this CTA = row j of its work item i / BH; (item, j) advance incrementally (no divisions per row)
        typedef EpiVec8T<(LEAN && MSB_EPI_X2) ? 1 : 3, !(LEAN && MSB_EPI_X2)> Ops;
        Ops ops;
        const int grid = gridDim.x;
        int j = g, item = blockIdx.x;
        int n = item / bands_per_img;
        int h = (item - n * bands_per_img) * BH + j;
        if (g < my_rows) epi_prefetch_vec8(epi, (((size_t)n * H + h) * W + rp) * C + 8 * k8, ops);
        for (int i = g; i < my_rows; i += 4) {
            // coordinates of this group's next row (i + 4)
            int j2 = j + 4, n2 = n, h2 = h + 4;
            if (j2 >= BH) {
                j2 -= BH; item += grid;
                n2 = item / bands_per_img;
                h2 = (item - n2 * bands_per_img) * BH + j2;
            }
            const int acc = i % kAccBufs;
            const uint32_t full_ph = (uint32_t)(i / kFullBars) & 1u;
            const EpiCoef coef = epi_coef(epi, n);
#ifdef MSB_CONV_DEBUG
#define TCT_ROW i
            TCT_WAIT(&bars->tmem_full[i % kFullBars], full_ph, 3);
#undef TCT_ROW
#else
            ptx::mbar_wait_backoff(&bars->tmem_full[i % kFullBars], full_ph, backoff_ns);
#endif
            ptx::tc_fence_after();
            if (dbg & 1) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
                j = j2; n = n2; h = h2;
                continue;
            }
            // ---- drain: hi + lo activation planes (columns j / 32 + j), then hi + lo weight rows (lanes l / l ^ 16) ----
            float o[16];
            {
                const uint32_t t_acc = tmem_base + (uint32_t)acc * kAccCols + lane_addr;
                float s0[16], s1[16];
                if (P3) {                     // the tensor core has already summed the two activation planes
                    ptx::tmem_ld16(t_acc + 0, s0);
                    ptx::tmem_ld16(t_acc + 16, s1);
                    ptx::tmem_ld_wait();
                } else {
                    float a[16], b[16];
                    ptx::tmem_ld16(t_acc + 0, a);
                    ptx::tmem_ld16(t_acc + 32, b);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) s0[j] = a[j] + b[j];
                    ptx::tmem_ld16(t_acc + 16, a);
                    ptx::tmem_ld16(t_acc + 48, b);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) s1[j] = a[j] + b[j];
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);     // accumulator back to the MMA warp
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float send = up ? s0[j] : s1[j];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
                    const float mine = up ? s1[j] : s0[j];
                    o[j] = mine + recv;                   // (hi row) + (lo row) in either lane: commutative, bit-identical
                }
            }
            named_bar_sync(1 + g, 128);                   // every thread of the group is done reading the previous row
#pragma unroll
            for (int j = 0; j < 16; ++j)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_addr + (uint32_t)j * 256u), "f"(o[j]) : "memory");
            named_bar_sync(1 + g, 128);                   // the row is staged
            const size_t row0 = (size_t)n * H + h;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int px = it * 16 + rp;
                float v[8];
                {
                    const int sw = (it << 4) ^ ((k8 >> 2) << 2);
                    const uint32_t a0 = stage + 4u * (uint32_t)(px * 64 + ((8 * k8) ^ sw));
                    const uint32_t a1 = stage + 4u * (uint32_t)(px * 64 + ((8 * k8 + 4) ^ sw));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a0));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a1));
                }
                const size_t idx = (row0 * W + px) * C + 8 * k8;
                const size_t sidx = (row0 * 2) * plane_stride + (size_t)px * C + 8 * k8;
                // the operand loads of the next 8-channel group are issued from inside the finish, as soon as this group's
                // operands are consumed: their latency hides behind this group's activation / split arithmetic
                const bool has_next = it == 0 || i + 4 < my_rows;
                const size_t nidx = it == 0 ? (row0 * W + 16 + rp) * C + 8 * k8 : (((size_t)n2 * H + h2) * W + rp) * C + 8 * k8;
                epi_finish_v8<ACT>(epi, coef, v, ops, idx, sidx, plane_stride,
                                   [&]() { if (has_next) epi_prefetch_vec8(epi, nidx, ops); });
            }
            j = j2; n = n2; h = h2;
        }
    }
    if (warp >= kMmaWarp0) TCT_STAMP(3);
    if (warp < kEpiWarps) TCT_STAMP(4);
    ptx::tc_fence_before();
    __syncthreads();
#ifdef MSB_CONV_DEBUG
    if (threadIdx.x == 0 && my_rows > 0) {
        const unsigned long long t0 = dbg_t[0];
        for (int i = 1; i <= 4; ++i) atomicAdd(&g_tct_time[i - 1], dbg_t[i] > t0 ? dbg_t[i] - t0 : 0ull);
        atomicAdd(&g_tct_time[4], (unsigned long long)clock64() - t0);
        atomicAdd(&g_tct_time[5], 1ull);
    }
#endif
    if (warp == kAllocWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int ACT, bool LEAN, bool P3>
int launch_act(const __nv_bfloat16* split_in, const void* wpacked, const EpiParams& epi, ConvShape s, cudaStream_t st) {
    CUtensorMap tm_act;
    if (make_tmap_split5d(&tm_act, split_in, s.B, s.H, s.W, s.C, 32, 1)) return -1;
    constexpr size_t smem = (size_t)kSlots * kSlotBytes + 4 * kStageBytes + sizeof(Barriers) + 1024;
    static_assert(smem <= 232448, "shared memory per CTA");
    auto kern = conv3x3_tct_kernel<ACT, LEAN, P3>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(conv3x3_tct)"))
        return -1;
    int BH = tune_get(TUNE_TCT_BAND);
    if (BH < 4 || s.H % BH) BH = s.H % 16 == 0 ? 16 : (s.H % 8 == 0 ? 8 : 4);       // >= 4: one epilogue group per row of a quad
    const int bands_per_img = s.H / BH;
    const int num_items = s.B * bands_per_img;
    const int grid = std::min(num_items, num_sms());
    const cudaError_t le = launch_maybe_pdl(kern, grid, kThreads, smem, st, tm_act, (const uint4*)wpacked, epi, s.H, num_items,
                                            bands_per_img, BH, tune_get(TUNE_EPI_L2_PREFETCH),
                                            (uint32_t)tune_get(TUNE_WAIT_BACKOFF), tune_get(TUNE_TCT_DEBUG));
    count_launch();
    return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "conv3x3_tct launch");
}

}  // namespace

int launch_conv3x3_tct(const __nv_bfloat16* split_in, const void* wpacked, const EpiParams& epi, ConvShape s, cudaStream_t st) {
    if (!tct_shape_supported(s.C, s.H, s.W)) {
        set_error("tcgen05 conv (TMEM-resident weights): unsupported shape C=%d H=%d W=%d", s.C, s.H, s.W);
        return -1;
    }
    if (epi.chan_bias || epi.pix_bias) { set_error("tcgen05 conv: bias terms are SIMT-engine only"); return -1; }
    const int act = (epi.out_split || epi.dact_out) ? epi.act : ACT_NONE;
    // lean instantiation: at most one src[] operand and no split_mul (every RK2 / Euler launch of the pre-activation RHS)
    const bool lean = epi_is_lean(epi);
    const bool p3 = tune_get(TUNE_TCT_PRODUCTS) == 3;
#define MSB_TCT_GO(A) (lean ? (p3 ? launch_act<A, true, true>(split_in, wpacked, epi, s, st) : launch_act<A, true, false>(split_in, wpacked, epi, s, st)) \
                            : (p3 ? launch_act<A, false, true>(split_in, wpacked, epi, s, st) : launch_act<A, false, false>(split_in, wpacked, epi, s, st)))
    if (act == ACT_GELU) return MSB_TCT_GO(ACT_GELU);
    if (act == ACT_RELU) return MSB_TCT_GO(ACT_RELU);
    return MSB_TCT_GO(ACT_NONE);
#undef MSB_TCT_GO
}

int tct_products() { return tune_get(TUNE_TCT_PRODUCTS) == 3 ? 3 : 4; }

// ---------------------------------------------------------------------------------------------
// weight pack for this form: OIHW fp32 -> the TMEM image of A = [W_hi ; W_lo], as 32-bit words
//   out[chunk (18)][lane L (128)][16]      column = chunk * 16 + j = tap * 32 + c_in / 2,
//   word = bf16(c_in even) | bf16(c_in odd) << 16,
//   lane L: quadrant q = L / 32, l = L % 32: hi plane of output channel 16 q + l (l < 16), lo plane of 16 q + l - 16
//   (the two rows of one channel sit 16 lanes apart inside a quadrant: one shuffle adds them in the epilogue).
// transpose = the input-gradient convolution (W^T, rotated 180 degrees).
// ---------------------------------------------------------------------------------------------
__global__ void pack_w_tct_kernel(const float* __restrict__ w, uint32_t* __restrict__ out, int transpose, int cin_total, int skip_in) {
    constexpr int C = 64;
    const int total = 128 * 288;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int j = i & 15;
        const int L = (i >> 4) & 127;
        const int chunk = i >> 11;
        const int col = chunk * 16 + j;
        const int tap = col >> 5, ci = (col & 31) * 2;
        const int l = L & 31, q = L >> 5;
        const int part = l >= 16, co = 16 * q + (l & 15);
        const int r = tap / 3, s = tap % 3;
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (!transpose) v[e] = w[(((size_t)co * cin_total + (ci + e) + skip_in) * 3 + r) * 3 + s];
            else v[e] = w[(((size_t)(ci + e) * cin_total + co + skip_in) * 3 + (2 - r)) * 3 + (2 - s)];
        }
        __nv_bfloat16 hi0, lo0, hi1, lo1;
        split_bf16(v[0], hi0, lo0);
        split_bf16(v[1], hi1, lo1);
        const uint32_t e0 = __bfloat16_as_ushort(part ? lo0 : hi0), e1 = __bfloat16_as_ushort(part ? lo1 : hi1);
        out[i] = e0 | (e1 << 16);
    }
}

// cin_total / skip_in: see pack_w_tc_kernel (MNIST ConcatConv2d weights carry the time channel first)
void launch_pack_w_tct_ex(const float* w, void* out, int transpose, int cin_total, int skip_in, cudaStream_t st) {
    const int total = 128 * 288;
    pack_w_tct_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, (uint32_t*)out, transpose, cin_total, skip_in);
    count_launch();
}
void launch_pack_w_tct(const float* w, void* out, int transpose, cudaStream_t st) { launch_pack_w_tct_ex(w, out, transpose, 64, 0, st); }

}  // namespace msb
