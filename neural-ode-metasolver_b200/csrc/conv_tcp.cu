// tcgen05 engine, pixel-major form: 3x3 (stride 1, pad 1) convolution as an implicit GEMM with the OUTPUT
// PIXELS on the M side and the output channels on the N side.
//
//   D[pixel][c_out] = sum_{tap, c_in} X[pixel + tap][c_in] * W[c_out][tap][c_in]
//
// GEMM roles:  A (M = 128 rows) = activations: 128 pixels (ROWS full image rows of one image) of ONE bf16
//                                 plane (hi or lo) of the split tensor, K-major (channels contiguous)
//              B (N rows)       = weights, K-major tiles [W_hi (C rows) ; W_lo (C rows)] per (tap, 64-wide c_in chunk)
//              D                = 128 TMEM lanes (pixels) x 2C fp32 columns: [X * W_hi | X_hi * W_lo], double buffered
// Precision: operands are bf16 hi + bf16 lo.  Three of the four hi/lo products are formed,
//      X_hi*W_hi, X_hi*W_lo  (one N = 2C MMA)   and   X_lo*W_hi  (one N = C MMA into the W_hi columns);
// the dropped X_lo*W_lo term is <= 2^-18 relative per product (typically 4e-7 of the result: below the fp32
// accumulation noise, measured in tests/).  25 % fewer tensor-core cycles than forming all four.
//
// Why pixels on M: a TMEM lane is then a pixel, so an epilogue thread owns ONE pixel and walks its channels
// -- NHWC-contiguous.  All epilogue traffic is whole 32-byte sectors (256-bit loads/stores), bf16 pairs are
// converted packed, addresses are formed once per 8 elements: ~3x fewer issued instructions per output
// element than the channel-major form (conv_tc.cu), which was issue-bound in its epilogue (profiles/).
//
// Data movement: per (c_in chunk, horizontal tap s) two TMA boxes (one per plane) bring ROWS+2 image rows x
// W pixels x 64 channels into shared memory as [plane][row][pixel] (zero-filled outside the image = the
// padding); the three vertical taps r are row offsets of the A descriptor.  Weight tiles stream through
// their own ring.  Warp roles: 0 = activation TMA, 1 = MMA issue, 2 = TMEM alloc, 3 = weight TMA,
// 4..11 = epilogue.  Persistent: grid = min(#tiles, #SMs).
#include <cuda.h>

#include "msb_internal.h"
#include "msb_ptx.cuh"

namespace msb {

int make_tmap_split5d(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h);
int make_tmap_rows64(CUtensorMap* m, const void* base, size_t rows, int box_rows);
// split tensor [B][H][2][W][C] bf16, one plane per box: box = {64 ch, box_w, 1, box_h, 1}, 128B swizzle, zero OOB fill
int make_tmap_split_plane(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h);

size_t tcp_packed_weight_bytes(int C) { return (size_t)9 * (C / 64) * (2 * C) * 64 * 2; }

#ifdef MSB_CONV_DEBUG
int tcp_debug_set(int flags) { return cudaMemcpyToSymbol(g_conv_debug, &flags, sizeof(int)) == cudaSuccess ? 0 : -1; }
int tcp_debug_hint(unsigned ns) { return cudaMemcpyToSymbol(ptx::g_suspend_hint, &ns, sizeof(ns)) == cudaSuccess ? 0 : -1; }
#endif

namespace {

constexpr int kEpiWarp0 = 4;
// EW = epilogue warps, 8 or 16.  With 8 (2 per scheduler) each thread software-pipelines its operand loads one chunk
// ahead in registers (2 x 48 registers of buffers; 160 registers / thread).  With 16 (4 per scheduler, 640 threads,
// <= 102 registers) the loads are issued at the top of their own chunk instead and the latency is covered by the other
// warps of the scheduler plus the L2 prefetch of the operands (the epilogue at 8 warps was latency-, not issue-bound:
// 29 % issue utilisation, profiles/).
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 212992;        // 8 epilogue warps: operand rings only
constexpr int kSmemBudget16 = 230400;      // 16 epilogue warps: rings + 64 KB of transpose staging (227 KB max per CTA)
constexpr int kStageBytesPerWarp = 4096;   // 32 pixels x 32 channels fp32

template <int C, int WIMG, int EW = 8> struct Geom {
    static constexpr int ROWS = 128 / WIMG;                          // image rows per 128-pixel tile
    static constexpr int PLANE_BYTES = (ROWS + 2) * WIMG * 128;      // one plane of the halo box
    static constexpr int X_STAGE_BYTES = 2 * PLANE_BYTES;
    static constexpr int ROW_BYTES = WIMG * 128;                     // one image row of one plane
    static constexpr int W_TILE_BYTES = 2 * C * 128;                 // [W_hi ; W_lo] x 64 k
    static constexpr int X_STAGES = 2;
    static constexpr int STAGE_BYTES = (EW == 16) ? 16 * kStageBytesPerWarp : 0;
    static constexpr int RING_BUDGET = (EW == 16 ? kSmemBudget16 : kSmemBudget) - STAGE_BYTES - X_STAGES * X_STAGE_BYTES;
    static constexpr int W_STAGES = RING_BUDGET / W_TILE_BYTES > kMaxStages ? kMaxStages : RING_BUDGET / W_TILE_BYTES;
    static constexpr int ACC_COLS = 2 * C;                           // fp32 accumulator columns per tile
    static constexpr int ACC_BUFS = (C == 64) ? 4 : 2;
    static_assert(ACC_COLS * ACC_BUFS <= 512, "TMEM");
    static_assert(W_STAGES >= 2, "weight ring");
};

struct __align__(8) Barriers {
    uint64_t w_full[kMaxStages], w_empty[kMaxStages];
    uint64_t x_full[kMaxStages], x_empty[kMaxStages];
    uint64_t tmem_full[4], tmem_empty[4];
    uint32_t tmem_base;
};

template <int C, int WIMG, int ACT, int EW>
__global__ void __launch_bounds__((kEpiWarp0 + EW) * 32, 1)
conv3x3_tcp_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                   const EpiParams epi, const int H, const int num_tiles, const int tiles_per_img, const int l2pf_dist,
                   const uint32_t backoff_ns, const int role_shift) {
    using G = Geom<C, WIMG, EW>;
    constexpr int CHUNKS = C / 64;
    constexpr int kWStages = G::W_STAGES, kXStages = G::X_STAGES, kAccBufs = G::ACC_BUFS;
    constexpr int kEpiWarps = EW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_x = smem;                                          // kXStages x X_STAGE_BYTES
    uint8_t* smem_w = smem + kXStages * G::X_STAGE_BYTES;            // kWStages x W_TILE_BYTES
    uint8_t* smem_stage = smem_w + kWStages * G::W_TILE_BYTES;       // EW == 16: 4 KB per epilogue warp
    Barriers* bars = reinterpret_cast<Barriers*>(smem_stage + G::STAGE_BYTES);

    // role_shift = kEpiWarp0 (option mma_warp_high, default) rotates the roles so that the epilogue warps are the LOW
    // physical warps and TMA / MMA / alloc the last four: the schedulers favour the highest warp ids of a sub-partition,
    // and a late MMA issue is a tensor-pipe bubble while a late epilogue instruction is not.  (physical + 4) keeps
    // warp & 3, the TMEM lane quadrant a warp may read.
    const int warp = (int)(((threadIdx.x >> 5) + role_shift) % (kEpiWarp0 + EW));
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        ptx::prefetch_tmap(&tmap_w);
        for (int i = 0; i < kWStages; ++i) { ptx::mbar_init(&bars->w_full[i], 1); ptx::mbar_init(&bars->w_empty[i], 1); }
        for (int i = 0; i < kXStages; ++i) { ptx::mbar_init(&bars->x_full[i], 1); ptx::mbar_init(&bars->x_empty[i], 1); }
        for (int i = 0; i < kAccBufs; ++i) { ptx::mbar_init(&bars->tmem_full[i], 1); ptx::mbar_init(&bars->tmem_empty[i], EW == 16 ? 4 * (C / 32) : kEpiWarps); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(&bars->tmem_base, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    ptx::pdl_launch_dependents();       // the next kernel may set itself up while this one runs ...
    ptx::pdl_wait();                    // ... and this one touches its inputs only after its predecessor has finished

    if (warp == 0) {
        // ===================== activation producer =====================
        if (lane == 0) {
            // bulk L2 prefetch of the epilogue's operands, l2pf_dist tiles ahead (see conv_tc.cu); a tile is
            // ROWS full image rows = one contiguous 128 x C fp32 range of every operand
            const bool pf_n = l2pf_dist > 0;
            auto l2_prefetch_tile = [&](int t) {
                if (t >= num_tiles) return;
                const int n = t / tiles_per_img;
                const int h0 = (t - n * tiles_per_img) * G::ROWS;
#pragma unroll
                for (int i = 0; i < kEpiLoadSlots; ++i) {
                    const float* p = epi_load_operand(epi, i);
                    if (p) ptx::l2_prefetch_bulk(p + ((size_t)n * H + h0) * WIMG * C, 128 * C * 4);
                }
            };
            if (pf_n)
                for (int d = 0; d + 1 < l2pf_dist; ++d) l2_prefetch_tile(blockIdx.x + d * gridDim.x);
            int st = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_img;
                const int h0 = (tile - n * tiles_per_img) * G::ROWS;
                if (pf_n) l2_prefetch_tile(tile + (l2pf_dist - 1) * gridDim.x);
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s) {
                        ptx::mbar_wait(&bars->x_empty[st], ph ^ 1);
                        if (MSB_DBG(8)) { ptx::mbar_arrive(&bars->x_full[st]); }
                        else {
                        ptx::mbar_arrive_expect_tx(&bars->x_full[st], G::X_STAGE_BYTES);
                        uint8_t* dst = smem_x + st * G::X_STAGE_BYTES;
                        ptx::tma_load_5d(dst, &tmap_act, &bars->x_full[st], chunk * 64, s - 1, 0, h0 - 1, n);
                        ptx::tma_load_5d(dst + G::PLANE_BYTES, &tmap_act, &bars->x_full[st], chunk * 64, s - 1, 1, h0 - 1, n);
                        }
                        if (++st == kXStages) { st = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer =====================
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s)
                        for (int r = 0; r < 3; ++r) {
                            const int wt = (r * 3 + s) * CHUNKS + chunk;
                            ptx::mbar_wait(&bars->w_empty[st], ph ^ 1);
                            if (MSB_DBG(4)) { ptx::mbar_arrive(&bars->w_full[st]); }
                            else {
                            ptx::mbar_arrive_expect_tx(&bars->w_full[st], G::W_TILE_BYTES);
                            ptx::tma_load_2d(smem_w + st * G::W_TILE_BYTES, &tmap_w, &bars->w_full[st], 0, wt * 2 * C);
                            }
                            if (++st == kWStages) { st = 0; ph ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc_full = ptx::make_idesc_bf16(128, 2 * C, 0, 0);   // X_hi * [W_hi ; W_lo]
            constexpr uint32_t idesc_hi = ptx::make_idesc_bf16(128, C, 0, 0);         // X_lo * W_hi
            int wst = 0, xst = 0; uint32_t wph = 0, xph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G::ACC_COLS);
                uint32_t accumulate = 0;
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s) {
                        ptx::mbar_wait(&bars->x_full[xst], xph);
                        ptx::tc_fence_after();
                        const uint32_t x_base = ptx::smem_u32(smem_x + xst * G::X_STAGE_BYTES);
                        for (int r = 0; r < 3; ++r) {
                            ptx::mbar_wait(&bars->w_full[wst], wph);
                            ptx::tc_fence_after();
                            const uint32_t w_base = ptx::smem_u32(smem_w + wst * G::W_TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t wdesc = ptx::make_smem_desc_sw128(w_base + k * 32, 16, 1024);
                                const uint64_t xhi = ptx::make_smem_desc_sw128(x_base + r * G::ROW_BYTES + k * 32, 16, 1024);
                                const uint64_t xlo =
                                    ptx::make_smem_desc_sw128(x_base + G::PLANE_BYTES + r * G::ROW_BYTES + k * 32, 16, 1024);
                                if (!MSB_DBG(2)) {
                                ptx::umma_bf16(d_tmem, xhi, wdesc, idesc_full, accumulate);
                                ptx::umma_bf16(d_tmem, xlo, wdesc, idesc_hi, 1u);
                                }
                                accumulate = 1;
                            }
                            ptx::umma_commit(&bars->w_empty[wst]);
                            if (++wst == kWStages) { wst = 0; wph ^= 1; }
                        }
                        ptx::umma_commit(&bars->x_empty[xst]);
                        if (++xst == kXStages) { xst = 0; xph ^= 1; }
                    }
                ptx::umma_commit(&bars->tmem_full[acc]);
                if (++acc == kAccBufs) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue =====================
        // thread = one pixel (TMEM lane) x a contiguous range of channels, 8 channels per chunk.  Operand
        // loads of chunk i+1 (or of chunk 0 of the next tile, before waiting for its accumulator) are in
        // flight while chunk i is computed and stored.
        const int we = warp - kEpiWarp0;
        const int q = warp & 3;                               // TMEM lane quadrant this warp may read
        constexpr int NGROUPS = kEpiWarps / 4;
        constexpr int CG = C / NGROUPS;                       // channels per warp group
        constexpr int NCHUNK = CG / 8;
        static_assert(NCHUNK % 2 == 0, "chunk pipeline is unrolled by two");
        const int cbase = (we >> 2) * CG;
        uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        asm volatile("" : "+r"(lane_addr));
        const int p = q * 32 + lane;                          // pixel of the tile
        const int rr = p / WIMG, wq = p - rr * WIMG;
        const size_t plane_stride = (size_t)WIMG * C;

        int acc = 0; uint32_t acc_ph = 0;
        int tile = blockIdx.x;
        auto pix_of = [&](int t) {
            const int n = t / tiles_per_img, h0 = (t - n * tiles_per_img) * G::ROWS;
            return ((size_t)n * H + h0 + rr) * WIMG + wq;
        };
        if (EW == 16) {
            // ---- 16 warps, transposed ownership ----------------------------------------------------------------
            // A TMEM lane is a pixel, so a thread reading its own accumulator row owns 32 B of a pixel and the 32
            // lanes of a warp-wide global access touch 32 different 128-byte lines: measured (profiles/
            // power_probe_r1*.log) the epilogue's stores then cost 45-70 us per launch EVEN WHEN THEY HIT L2 --
            // the SM's load/store path (one line per cycle), not DRAM, was the limiter.  Here a warp (32 pixels x
            // 32 channels) passes its accumulators through a private, XOR-swizzled 4 KB shared-memory stage and
            // reads them back transposed within each quad of lanes: lane 4i+k then owns channels 8k..8k+7 of the
            // four pixels 4i..4i+3, so the four lanes of a quad cover one whole 128-byte line per access (8 lines
            // instead of 32 per warp instruction, for loads and stores alike).  The accumulator buffer is handed
            // back to the MMA warp as soon as it is copied out.  C = 64: 8 warps per tile, the two groups of 8
            // take alternate tiles; C = 128: all 16 warps on every tile.
            constexpr int NG = C / 32;                 // channel groups of 32 per tile
            constexpr int WPT = 4 * NG;                // warps per tile
            constexpr int TG = 16 / WPT;               // tile groups
            static_assert(TG >= 1 && 16 % WPT == 0, "16 epilogue warps must split evenly");
            const int tg = we / WPT, wi = we % WPT;
            const int cb = (wi >> 2) * 32;             // first channel of this warp's group
            const uint32_t stage = ptx::smem_u32(smem_stage + we * kStageBytesPerWarp);
            const int k4 = lane & 3, quad = lane >> 2;
            // owned element block j (j = 0..3): pixel 32q + 4*quad + j of the tile, channels cb + 8*k4 .. +7
            // (row0 = first image row of the tile counted over the whole batch: n * H + h0)
            auto owned = [&](size_t row0, int j, size_t& idx, size_t& sidx) {
                const int pp = q * 32 + quad * 4 + j;
                const int r2 = pp / WIMG, w2 = pp - r2 * WIMG;
                idx = ((row0 + r2) * WIMG + w2) * C + cb + 8 * k4;
                sidx = ((row0 + r2) * 2) * plane_stride + (size_t)w2 * C + cb + 8 * k4;
            };
            auto row0_of = [&](int t) {
                const int n = t / tiles_per_img;
                return (size_t)n * H + (size_t)(t - n * tiles_per_img) * G::ROWS;
            };
            int it = tg;
            tile = blockIdx.x + tg * gridDim.x;
            EpiVec8 ops;
            if (tile < num_tiles) { size_t i0, s0; owned(row0_of(tile), 0, i0, s0); epi_prefetch_vec8(epi, i0, ops); }
            for (; tile < num_tiles; tile += TG * gridDim.x, it += TG) {
                const int acc16 = it % kAccBufs;
                const uint32_t ph16 = (uint32_t)(it / kAccBufs) & 1u;
                const EpiCoef coef = epi_coef(epi, tile / tiles_per_img);
                const size_t row0 = row0_of(tile);
                ptx::mbar_wait_backoff(&bars->tmem_full[acc16], ph16, backoff_ns);
                ptx::tc_fence_after();
                if (MSB_DBG(1)) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc16]);
                    continue;
                }
                const uint32_t t_acc = tmem_base + (uint32_t)(acc16 * G::ACC_COLS) + lane_addr + (uint32_t)cb;
                __syncwarp();                          // every lane is done reading the previous tile's stage
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {       // 8 channels = two 16-byte units of this lane's 128-byte row
                    float a[8], b[8];
                    ptx::tmem_ld<8>(t_acc + c8 * 8, a);
                    ptx::tmem_ld<8>(t_acc + C + c8 * 8, b);
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
                if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc16]);
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
                    size_t idx, sidx;
                    owned(row0, j, idx, sidx);
                    epi_finish_v8<ACT>(epi, coef, v, ops, idx, sidx, plane_stride);
                    if (j < 3) {
                        owned(row0, j + 1, idx, sidx);
                        epi_prefetch_vec8(epi, idx, ops);
                    } else {
                        const int t2 = tile + TG * gridDim.x;
                        if (t2 < num_tiles) { owned(row0_of(t2), 0, idx, sidx); epi_prefetch_vec8(epi, idx, ops); }
                    }
                }
            }
        } else {
        EpiVec8 opsA, opsB;
        if (tile < num_tiles) epi_prefetch_vec8(epi, pix_of(tile) * C + cbase, opsA);
        for (; tile < num_tiles; tile += gridDim.x) {
            const int n = tile / tiles_per_img;
            const int h = (tile - n * tiles_per_img) * G::ROWS + rr;
            const size_t pix = ((size_t)n * H + h) * WIMG + wq;
            const size_t idx_t = pix * C + cbase;
            const size_t split_t = (((size_t)n * H + h) * 2) * plane_stride + (size_t)wq * C + cbase;
            const EpiCoef coef = epi_coef(epi, n);
            ptx::mbar_wait_backoff(&bars->tmem_full[acc], acc_ph, backoff_ns);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)(acc * G::ACC_COLS) + lane_addr + (uint32_t)cbase;
            if (MSB_DBG(1)) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
                if (++acc == kAccBufs) { acc = 0; acc_ph ^= 1; }
                continue;
            }
            auto do_chunk = [&](const int ch, const EpiVec8& cur, EpiVec8& nxt) {
                float v[8];
                {
                    float a[8], b[8];
                    ptx::tmem_ld<8>(t_acc + ch * 8, a);
                    ptx::tmem_ld<8>(t_acc + C + ch * 8, b);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = a[j] + b[j];
                }
                if (ch == NCHUNK - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
                    const int t2 = tile + gridDim.x;
                    if (t2 < num_tiles) epi_prefetch_vec8(epi, pix_of(t2) * C + cbase, nxt);
                } else {
                    epi_prefetch_vec8(epi, idx_t + (ch + 1) * 8, nxt);
                }
                epi_finish_v8<ACT>(epi, coef, v, cur, idx_t + ch * 8, split_t + ch * 8, plane_stride);
            };
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ch += 2) {
                do_chunk(ch, opsA, opsB);
                do_chunk(ch + 1, opsB, opsA);
            }
            if (++acc == kAccBufs) { acc = 0; acc_ph ^= 1; }
        }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int C, int WIMG, int ACT, int EW>
int launch_act_ew(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
               cudaStream_t st) {
    using G = Geom<C, WIMG, EW>;
    CUtensorMap tm_act, tm_w;
    if (make_tmap_split_plane(&tm_act, split_in, s.B, s.H, s.W, s.C, WIMG, G::ROWS + 2)) return -1;
    if (make_tmap_rows64(&tm_w, w_tiles, tcp_packed_weight_bytes(C) / 128, 2 * C)) return -1;
    const size_t smem = (size_t)G::X_STAGES * G::X_STAGE_BYTES + (size_t)G::W_STAGES * G::W_TILE_BYTES + G::STAGE_BYTES +
                        sizeof(Barriers) + 1024;
    static_assert(G::X_STAGES * G::X_STAGE_BYTES + G::W_STAGES * G::W_TILE_BYTES + G::STAGE_BYTES + sizeof(Barriers) + 1024 <= 232448,
                  "shared memory per CTA");
    auto kern = conv3x3_tcp_kernel<C, WIMG, ACT, EW>;
    constexpr int kThreads = (kEpiWarp0 + EW) * 32;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(conv3x3_tcp)"))
        return -1;
    const int tiles_per_img = s.H / G::ROWS;
    const int num_tiles = s.B * tiles_per_img;
    const int grid = std::min(num_tiles, num_sms());
    const cudaError_t le = launch_maybe_pdl(kern, grid, kThreads, smem, st, tm_act, tm_w, epi, s.H, num_tiles, tiles_per_img,
                                            tune_get(TUNE_EPI_L2_PREFETCH), (uint32_t)tune_get(TUNE_WAIT_BACKOFF), tune_get(TUNE_MMA_WARP_HIGH) ? kEpiWarp0 : 0);
    count_launch();
    return check_cuda(le != cudaSuccess ? le : cudaGetLastError(), "conv3x3_tcp launch");
}

template <int C, int WIMG, int ACT>
int launch_act(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
               cudaStream_t st) {
    if (tune_get(TUNE_TCP_EPI_WARPS) == 16) return launch_act_ew<C, WIMG, ACT, 16>(split_in, w_tiles, epi, s, st);
    return launch_act_ew<C, WIMG, ACT, 8>(split_in, w_tiles, epi, s, st);
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

int launch_conv3x3_tcp(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
                       cudaStream_t st) {
    if (!tc_shape_supported(s.C, s.H, s.W)) {
        set_error("tcgen05 conv: unsupported shape C=%d H=%d W=%d", s.C, s.H, s.W);
        return -1;
    }
    if (epi.chan_bias || epi.pix_bias) { set_error("tcgen05 conv: bias terms are SIMT-engine only"); return -1; }
    if (s.C == 64 && s.W == 32) return launch_impl<64, 32>(split_in, w_tiles, epi, s, st);
    if (s.C == 64 && s.W == 16) return launch_impl<64, 16>(split_in, w_tiles, epi, s, st);
    if (s.C == 128 && s.W == 32) return launch_impl<128, 32>(split_in, w_tiles, epi, s, st);
    return launch_impl<128, 16>(split_in, w_tiles, epi, s, st);
}

// ---------------------------------------------------------------------------------------------
// weight pack for this engine: OIHW fp32 -> bf16 tiles of 2C rows x 64 k (K-major, 128-byte rows; TMA applies
// the swizzle).  tile = tap * (C/64) + chunk; row n < C: hi(W[co = n]), row n >= C: lo(W[co = n - C]);
// element k = input channel chunk*64 + k.  transpose = the input-gradient convolution (W^T, rotated 180 degrees).
// ---------------------------------------------------------------------------------------------
__global__ void pack_w_tcp_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int C, int transpose) {
    const int chunks = C / 64;
    const int total = 9 * chunks * 2 * C * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i & 63;
        int rest = i >> 6;
        const int row = rest % (2 * C); rest /= 2 * C;
        const int chunk = rest % chunks;
        const int tap = rest / chunks;
        const int part = row >= C, co = part ? row - C : row;
        const int ci = chunk * 64 + k;
        const int r = tap / 3, s = tap % 3;
        float v;
        if (!transpose) v = w[(((size_t)co * C + ci) * 3 + r) * 3 + s];
        else v = w[(((size_t)ci * C + co) * 3 + (2 - r)) * 3 + (2 - s)];
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        out[i] = part ? lo : hi;
    }
}

void launch_pack_w_tcp(const float* w, __nv_bfloat16* out, int C, int transpose, cudaStream_t st) {
    const int total = 9 * (C / 64) * 2 * C * 64;
    pack_w_tcp_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, out, C, transpose);
    count_launch();
}

}  // namespace msb
