// tcgen05 engine: 3x3 (stride 1, pad 1) convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D^T[c_out][pixel] = sum_{tap, c_in} W[c_out][tap][c_in] * X[pixel + tap][c_in]
//
// GEMM roles:  A (M side, 128 rows)  = weights, K-major bf16 tiles packed by pack_w_tc_kernel
//              B (N side, 256 rows)  = activations straight out of the split tensor [B][H][2][W][C]:
//                                      128 pixels x {hi, lo} planes, K-major (channels contiguous)
//              D                     = 128 lanes x 256 fp32 columns in TMEM, double buffered (512 cols)
// Precision: every operand lives as bf16 hi + bf16 lo (~16 significand bits).  With [W_hi;W_lo] on
// the M side (C=64: two 64-row halves of one tile; C=128: two tiles accumulated into the same D)
// and [X_hi | X_lo] on the N side, one N=256 MMA chain forms all four hi/lo products in fp32; the
// epilogue adds the hi/lo columns (and, for C=64, the hi/lo rows with one warp shuffle).
// Measured against the reference: 4.6e-7 max-rel on ODE-block outputs == fp32-vs-fp64 noise.
//
// Data movement: one TMA 5-D box per (c_in chunk, horizontal tap s) brings ROWS+2 image rows x 2
// planes x W pixels x 64 channels (zero-filled outside the image = the conv padding) and serves the
// three vertical taps r by offsetting the smem descriptor by r image rows.  Weight tiles stream
// through their own ring.  Warp roles: 0 = activation TMA, 1 = MMA issue, 2 = TMEM alloc,
// 3 = weight TMA, 4..11 = epilogue (TMEM -> registers -> fused RK epilogue -> global).
// Persistent: grid = min(#tiles, #SMs); tile = 128 pixels (ROWS full image rows of one image).
#include <cuda.h>

#include "msb_internal.h"
#include "msb_ptx.cuh"

#ifndef MSB_B_STAGES
#define MSB_B_STAGES 2
#endif

namespace msb {

// ---------------------------------------------------------------------------------------------
// host: TMA descriptors
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
            return nullptr;
        }
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// split tensor [B][H][2][W][C] bf16, box = {64 ch, box_w, 2 planes, box_h, 1 image}, 128B swizzle, zero OOB fill
int make_tmap_split5d(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * 2 * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 2, (cuuint32_t)box_h, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(split 5d) failed with %d", (int)r); return -1; }
    return 0;
}

// same tensor, ONE plane per box: box = {64 ch, box_w, 1, box_h, 1} (pixel-major engine: smem = [row][pixel] per plane)
int make_tmap_split_plane(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * 2 * W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(split plane) failed with %d", (int)r); return -1; }
    return 0;
}

// row-major [rows][64] bf16 tiles, box = {64, box_rows}
int make_tmap_rows64(CUtensorMap* m, const void* base, size_t rows, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rows64) failed with %d", (int)r); return -1; }
    return 0;
}

#ifdef MSB_CONV_DEBUG
int tcp_debug_set(int flags);
int tct_debug_set(int flags);
int tcp_debug_hint(unsigned ns);
extern "C" int msb_debug_suspend_hint(unsigned ns) {
    if (tcp_debug_hint(ns)) return -1;
    return cudaMemcpyToSymbol(ptx::g_suspend_hint, &ns, sizeof(ns)) == cudaSuccess ? 0 : -1;
}
extern "C" int msb_debug_conv_flags(int flags) {
    if (tcp_debug_set(flags) || tct_debug_set(flags)) return -1;
    return cudaMemcpyToSymbol(g_conv_debug, &flags, sizeof(int)) == cudaSuccess ? 0 : -1;
}
#endif

bool tc_shape_supported(int C, int H, int W) {
    if (C != 64 && C != 128) return false;
    if (W == 32) return H % 4 == 0;
    if (W == 16) return H % 8 == 0;
    return false;
}
size_t tc_packed_weight_bytes(int C) { return (size_t)((C == 64) ? 9 : 9 * (C / 64) * 2) * 128 * 64 * 2; }

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kEpiWarp0 = 4;
constexpr int kATileBytes = 128 * 128;   // 128 rows x 64 bf16
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxStages = 8;

// Per-shape tuning.  C=64 has half the MMA time per output element of C=128, so its epilogue gets 16
// warps (4-pixel chunks) and the weight ring is deepened at the cost of one activation stage:
// the weight tiles turn over every 512 MMA cycles, far less than a loaded-L2 TMA round trip.
// RES: all weight tiles stay resident in shared memory for the whole kernel (C = 64 only: 9 x 16 KB), loaded once
// per CTA instead of once per tile -- the weight stream was half of the L2 -> SM traffic (profiles/).  To make room
// the tile is 8 rows x 16 pixels (two 40 KB activation stages) also on 32-pixel-wide images: WIMG is the TILE width,
// the image width is a runtime parameter (a multiple of WIMG).
template <int C, int WIMG, bool RES = false> struct TileGeom {
    static constexpr int ROWS = 128 / WIMG;                       // image rows per 128-pixel tile
    static constexpr int B_STAGE_BYTES = (ROWS + 2) * 2 * WIMG * 128;
    static constexpr int ROW_PAIR_BYTES = 2 * WIMG * 128;         // one image row, both planes
    static constexpr int EPI_WARPS = (C == 64) ? 16 : 8;          // multiple of 4 (one per TMEM lane quadrant)
    static constexpr int THREADS = (kEpiWarp0 + EPI_WARPS) * 32;
    static constexpr int PXO = (C == 64) ? 4 : 8;                 // pixels one epilogue thread owns per chunk
    static constexpr int B_STAGES = MSB_B_STAGES;
    static constexpr int A_RING = (212992 - B_STAGES * B_STAGE_BYTES) / kATileBytes > kMaxStages
                                      ? kMaxStages : (212992 - B_STAGES * B_STAGE_BYTES) / kATileBytes;
    static constexpr int A_STAGES = RES ? 9 : A_RING;             // smem weight tiles
    static_assert(!RES || C == 64, "resident weights: C = 64 only");
};

struct __align__(8) Barriers {
    uint64_t a_full[kMaxStages], a_empty[kMaxStages];
    uint64_t b_full[kMaxStages], b_empty[kMaxStages];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int C, int WIMG, int ACT, bool RES>
__global__ void __launch_bounds__((TileGeom<C, WIMG, RES>::THREADS), 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                  const EpiParams epi, const int H, const int Wimg, const int num_tiles, const int tiles_per_img,
                  const int l2pf_dist, const uint32_t backoff_ns) {
    using G = TileGeom<C, WIMG, RES>;
    const int cols = Wimg / WIMG;                                  // tile columns per image
    // tile -> (image n, first row h0, first pixel column w_off); neighbouring tiles are neighbouring columns
    auto tile_coords = [&](int tile, int& n, int& h0, int& w_off) {
        n = tile / tiles_per_img;
        const int t = tile - n * tiles_per_img;
        const int band = t / cols;
        h0 = band * G::ROWS;
        w_off = (t - band * cols) * WIMG;
    };
    constexpr int CHUNKS = C / 64;
    constexpr int kAStages = G::A_STAGES, kBStages = G::B_STAGES, kNumEpiWarps = G::EPI_WARPS, kPXO = G::PXO;
    constexpr int PARTS = (C == 64) ? 1 : 2;          // weight tiles per (tap, chunk): C=64 packs hi/lo into one tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem;                                        // kBStages x B_STAGE_BYTES
    uint8_t* smem_a = smem + kBStages * G::B_STAGE_BYTES;          // kAStages x 16 KB
    Barriers* bars = reinterpret_cast<Barriers*>(smem_a + kAStages * kATileBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        ptx::prefetch_tmap(&tmap_w);
        for (int i = 0; i < kAStages; ++i) { ptx::mbar_init(&bars->a_full[i], 1); ptx::mbar_init(&bars->a_empty[i], 1); }
        for (int i = 0; i < kBStages; ++i) { ptx::mbar_init(&bars->b_full[i], 1); ptx::mbar_init(&bars->b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars->tmem_full[i], 1); ptx::mbar_init(&bars->tmem_empty[i], kNumEpiWarps); }
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

    if (warp == 0) {
        // ===================== activation producer =====================
        if (lane == 0) {
            // The epilogue's operands (y, k_j, act' ... of the SAME pixels, fp32 NHWC) are pulled into L2 by bulk
            // prefetches issued here, l2pf_dist tiles ahead of the activation loads: the epilogue threads hold only
            // one chunk of loads in flight, which at HBM latency caps them far below the HBM bandwidth; from L2
            // the same loads are ~3x shorter (profiles/epi_l2_prefetch_r1.txt).
            const bool pf_n = l2pf_dist > 0;
            auto l2_prefetch_tile = [&](int t) {
                if (t >= num_tiles) return;
                int n, h0, w_off;
                tile_coords(t, n, h0, w_off);
#pragma unroll
                for (int i = 0; i < kEpiLoadSlots; ++i) {
                    const float* p = epi_load_operand(epi, i);
                    if (!p) continue;
                    if (cols == 1) {
                        ptx::l2_prefetch_bulk(p + ((size_t)n * H + h0) * Wimg * C, G::ROWS * WIMG * C * 4);
                    } else {
                        for (int rr = 0; rr < G::ROWS; ++rr)
                            ptx::l2_prefetch_bulk(p + (((size_t)n * H + h0 + rr) * Wimg + w_off) * C, WIMG * C * 4);
                    }
                }
            };
            if (pf_n)
                for (int d = 0; d + 1 < l2pf_dist; ++d) l2_prefetch_tile(blockIdx.x + d * gridDim.x);
            int st = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int n, h0, w_off;
                tile_coords(tile, n, h0, w_off);
                if (pf_n) l2_prefetch_tile(tile + (l2pf_dist - 1) * gridDim.x);
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s) {
                        ptx::mbar_wait(&bars->b_empty[st], ph ^ 1);
                        if (MSB_DBG(8)) { ptx::mbar_arrive(&bars->b_full[st]); }
                        else {
                        ptx::mbar_arrive_expect_tx(&bars->b_full[st], G::B_STAGE_BYTES);
                        ptx::tma_load_5d(smem_b + st * G::B_STAGE_BYTES, &tmap_act, &bars->b_full[st],
                                         chunk * 64, w_off + s - 1, 0, h0 - 1, n);
                        }
                        if (++st == kBStages) { st = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 3) {
        // ===================== weight producer =====================
        if (lane == 0 && RES) {
            // resident weights: all 9 tiles once, one barrier
            ptx::mbar_arrive_expect_tx(&bars->a_full[0], 9 * kATileBytes);
            for (int wt = 0; wt < 9; ++wt)
                ptx::tma_load_2d(smem_a + wt * kATileBytes, &tmap_w, &bars->a_full[0], 0, wt * 128);
        } else if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s)
                        for (int r = 0; r < 3; ++r)
                            for (int part = 0; part < PARTS; ++part) {
                                const int tap = r * 3 + s;
                                const int wt = (C == 64) ? tap : ((tap * CHUNKS + chunk) * 2 + part);
                                ptx::mbar_wait(&bars->a_empty[st], ph ^ 1);
                                if (MSB_DBG(4)) { ptx::mbar_arrive(&bars->a_full[st]); }
                                else {
                                ptx::mbar_arrive_expect_tx(&bars->a_full[st], kATileBytes);
                                ptx::tma_load_2d(smem_a + st * kATileBytes, &tmap_w, &bars->a_full[st], 0, wt * 128);
                                }
                                if (++st == kAStages) { st = 0; ph ^= 1; }
                            }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 256, 0, 0);
            int ast = 0, bst = 0; uint32_t aph = 0, bph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            if (RES) { ptx::mbar_wait(&bars->a_full[0], 0); ptx::tc_fence_after(); }
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                uint32_t accumulate = 0;
                for (int chunk = 0; chunk < CHUNKS; ++chunk)
                    for (int s = 0; s < 3; ++s) {
                        ptx::mbar_wait(&bars->b_full[bst], bph);
                        ptx::tc_fence_after();
                        const uint32_t b_base = ptx::smem_u32(smem_b + bst * G::B_STAGE_BYTES);
                        for (int r = 0; r < 3; ++r)
                            for (int part = 0; part < PARTS; ++part) {
                                if (RES) ast = r * 3 + s;
                                else { ptx::mbar_wait(&bars->a_full[ast], aph); ptx::tc_fence_after(); }
                                const uint32_t a_base = ptx::smem_u32(smem_a + ast * kATileBytes);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t adesc = ptx::make_smem_desc_sw128(a_base + k * 32, 16, 1024);
                                    const uint64_t bdesc =
                                        ptx::make_smem_desc_sw128(b_base + r * G::ROW_PAIR_BYTES + k * 32, 16, 1024);
                                    if (!MSB_DBG(2)) ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                                    accumulate = 1;
                                }
                                if (!RES) {
                                    ptx::umma_commit(&bars->a_empty[ast]);
                                    if (++ast == kAStages) { ast = 0; aph ^= 1; }
                                }
                            }
                        ptx::umma_commit(&bars->b_empty[bst]);
                        if (++bst == kBStages) { bst = 0; bph ^= 1; }
                    }
                ptx::umma_commit(&bars->tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue =====================
        // Each thread owns one output channel and walks its share of the tile in chunks of 8
        // pixels.  Operand loads for chunk i+1 (and for chunk 0 of the NEXT tile, before waiting
        // for its accumulator) are in flight while chunk i is computed and stored.
        const int we = warp - kEpiWarp0;
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int part = we >> 2;                     // which slice of the tile's image rows
        uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        asm volatile("" : "+r"(lane_addr));           // keep it in a register (else S2R + shifts are redone per chunk)
        constexpr int NPARTS = kNumEpiWarps / 4;
        static_assert(G::ROWS % NPARTS == 0, "image rows of a tile must split evenly over the epilogue warp groups");
        constexpr int RH = G::ROWS / NPARTS;          // image rows per warp group
        constexpr int PX_PER_LD = (C == 64) ? 2 * kPXO : kPXO;   // pixels covered by one TMEM load step
        constexpr int STEPS_PER_ROW = WIMG / PX_PER_LD;
        constexpr int NCHUNK = RH * STEPS_PER_ROW;    // chunks (of kPXO owned pixels) per tile per thread
        const int sel = (C == 64) ? (lane >> 4) : 0;  // C=64: lanes l / l^16 split the pixels of a step
        const int c = (C == 64) ? (16 * q + (lane & 15)) : (32 * q + lane);
        const size_t plane_stride = (size_t)Wimg * C;

        // image row and image column of owned pixel 0 of chunk `ch` in the tile at (h0, w_off)
        auto chunk_pos = [&](int h0, int w_off, int ch, int& h, int& w0) {
            const int rr = ch / STEPS_PER_ROW, stp = ch - rr * STEPS_PER_ROW;
            h = h0 + part * RH + rr;
            w0 = w_off + stp * PX_PER_LD + sel * kPXO;
        };
        EpiOperands<kPXO> opsA, opsB;
        int acc = 0; uint32_t acc_ph = 0;
        int tile = blockIdx.x;
        if (tile < num_tiles) {
            int n, h0, w_off;
            tile_coords(tile, n, h0, w_off);
            int h, w0; chunk_pos(h0, w_off, 0, h, w0);
            epi_prefetch<kPXO>(epi, (((size_t)n * H + h) * Wimg + w0) * C + c, C, opsA);
        }
        for (; tile < num_tiles; tile += gridDim.x) {
            int n, h0, w_off;
            tile_coords(tile, n, h0, w_off);
            ptx::mbar_wait_backoff(&bars->tmem_full[acc], acc_ph, backoff_ns);
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)acc * 256u + lane_addr;
            const EpiCoef coef = epi_coef(epi, n);          // one image per tile: the slice is tile-uniform
            if (MSB_DBG(1)) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                continue;
            }
            auto do_chunk = [&](const int ch, const EpiOperands<kPXO>& cur, EpiOperands<kPXO>& nxt) {
                int h, w0; chunk_pos(h0, w_off, ch, h, w0);
                const int rr = ch / STEPS_PER_ROW, stp = ch - rr * STEPS_PER_ROW;
                const int rho = part * RH + rr;
                // ---- accumulator: hi + lo columns (and hi + lo weight rows for C = 64) ----
                float v[kPXO];
                const uint32_t col = (uint32_t)(rho * 2 * WIMG + stp * PX_PER_LD);
                if (C == 64) {
                    float hi[2 * kPXO], lo[2 * kPXO];
                    ptx::tmem_ld<2 * kPXO>(t_acc + col, hi);
                    ptx::tmem_ld<2 * kPXO>(t_acc + col + WIMG, lo);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 2 * kPXO; ++j) {
                        const float a = hi[j] + lo[j];
                        const float o = __shfl_xor_sync(0xffffffffu, a, 16);
                        hi[j] = sel ? o + a : a + o;       // identical operand order in both lanes
                    }
#pragma unroll
                    for (int j = 0; j < kPXO; ++j) v[j] = sel ? hi[kPXO + j] : hi[j];
                } else {
                    float hi[kPXO], lo[kPXO];
                    ptx::tmem_ld<kPXO>(t_acc + col, hi);
                    ptx::tmem_ld<kPXO>(t_acc + col + WIMG, lo);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < kPXO; ++j) v[j] = hi[j] + lo[j];
                }
                if (ch == NCHUNK - 1) {
                    // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[acc]);
                }
                // ---- prefetch the operands of the next chunk (possibly of the next tile) ----
                {
                    int n2 = n, h02 = h0, woff2 = w_off, ch2 = ch + 1;
                    bool have = true;
                    if (ch == NCHUNK - 1) {
                        const int t2 = tile + gridDim.x;
                        have = t2 < num_tiles;
                        tile_coords(t2, n2, h02, woff2); ch2 = 0;
                    }
                    if (have) {
                        int h2, w2; chunk_pos(h02, woff2, ch2, h2, w2);
                        epi_prefetch<kPXO>(epi, (((size_t)n2 * H + h2) * Wimg + w2) * C + c, C, nxt);
                    }
                }
                // ---- fused RK epilogue on the 8 owned pixels ----
                const size_t pix = ((size_t)n * H + h) * Wimg + w0;
                const size_t split0 = (((size_t)n * H + h) * 2) * plane_stride + (size_t)w0 * C + c;
                epi_finish<kPXO, ACT>(epi, coef, v, cur, pix * C + c, C, split0, plane_stride);
            };
            static_assert(NCHUNK % 2 == 0, "chunk pipeline is unrolled by two");
#pragma unroll
            for (int ch = 0; ch < NCHUNK; ch += 2) {
                do_chunk(ch, opsA, opsB);
                do_chunk(ch + 1, opsB, opsA);
            }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int C, int WIMG, int ACT, bool RES>
int launch_act(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
               cudaStream_t st) {
    using G = TileGeom<C, WIMG, RES>;
    constexpr int kAStages = G::A_STAGES, kBStages = G::B_STAGES;
    CUtensorMap tm_act, tm_w;
    if (make_tmap_split5d(&tm_act, split_in, s.B, s.H, s.W, s.C, WIMG, G::ROWS + 2)) return -1;
    const size_t wrows = tc_packed_weight_bytes(C) / 128;
    if (make_tmap_rows64(&tm_w, w_tiles, wrows, 128)) return -1;
    const size_t smem = (size_t)kBStages * G::B_STAGE_BYTES + (size_t)kAStages * kATileBytes + sizeof(Barriers) + 1024;
    auto kern = conv3x3_tc_kernel<C, WIMG, ACT, RES>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(conv3x3_tc)"))
        return -1;
    const int tiles_per_img = (s.H / G::ROWS) * (s.W / WIMG);
    const int num_tiles = s.B * tiles_per_img;
    const int grid = std::min(num_tiles, num_sms());
    kern<<<grid, G::THREADS, smem, st>>>(tm_act, tm_w, epi, s.H, s.W, num_tiles, tiles_per_img,
                                         tune_get(TUNE_EPI_L2_PREFETCH), (uint32_t)tune_get(TUNE_WAIT_BACKOFF));
    count_launch();
    return check_cuda(cudaGetLastError(), "conv3x3_tc launch");
}

template <int C, int WIMG, bool RES = false>
int launch_impl(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
                cudaStream_t st) {
    // the activation only matters when the epilogue emits act(out) / act'(out)
    const int act = (epi.out_split || epi.dact_out) ? epi.act : ACT_NONE;
    if (act == ACT_GELU) return launch_act<C, WIMG, ACT_GELU, RES>(split_in, w_tiles, epi, s, st);
    if (act == ACT_RELU) return launch_act<C, WIMG, ACT_RELU, RES>(split_in, w_tiles, epi, s, st);
    return launch_act<C, WIMG, ACT_NONE, RES>(split_in, w_tiles, epi, s, st);
}

// MSB_TC_RESIDENT=1 selects the resident-weight variant.  Measured (profiles/conv_forms_r1.txt): it halves the
// L2 -> SM traffic and speeds the MMA + TMA side up by 17 % (110 vs 133 us without the epilogue), but the full
// kernel is unchanged within noise (155.6 vs 151.6 us per launch) -- the epilogue's HBM traffic, not the operand
// feed, is what the MMAs wait for.  Off by default; kept as the starting point for the round-2 epilogue work.
bool resident_weights_enabled() { return tune_get(TUNE_TC_RESIDENT) == 1; }

}  // namespace

int launch_conv3x3_tc(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi, ConvShape s,
                      cudaStream_t st) {
    if (!tc_shape_supported(s.C, s.H, s.W)) {
        set_error("tcgen05 conv: unsupported shape C=%d H=%d W=%d", s.C, s.H, s.W);
        return -1;
    }
    if (s.C == 64 && s.H % 8 == 0 && resident_weights_enabled())       // 8 x 16 tiles, weights resident in shared memory
        return launch_impl<64, 16, true>(split_in, w_tiles, epi, s, st);
    if (s.C == 64 && s.W == 32) return launch_impl<64, 32>(split_in, w_tiles, epi, s, st);
    if (s.C == 64 && s.W == 16) return launch_impl<64, 16>(split_in, w_tiles, epi, s, st);
    if (s.C == 128 && s.W == 32) return launch_impl<128, 32>(split_in, w_tiles, epi, s, st);
    return launch_impl<128, 16>(split_in, w_tiles, epi, s, st);
}

}  // namespace msb
