// Internal (non-ABI) declarations shared by the .cu files of libmetasolver_b200.so
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>

#include "msb_common.cuh"

namespace msb {

// ---- runtime helpers (api.cu) ----
int num_sms();
void count_launch(int n = 1);
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);   // 0 if ok, else sets error and returns -1

// tuning options: process-wide ints, set by msb_set_option() or the environment (MSB_EPI_L2_PREFETCH, MSB_TC_RESIDENT)
enum { TUNE_EPI_L2_PREFETCH = 0,   // conv epilogue operands: bulk L2 prefetch distance in tiles (0 = off)
       TUNE_TC_RESIDENT = 1,       // channel-major C=64 conv: weights resident in shared memory
       TUNE_TCP_EPI_WARPS = 2,     // pixel-major conv: epilogue warps (8 or 16)
       TUNE_TC_FORM_C64 = 3,       // tcgen05 conv form for C = 64: 0 = channel-major, 1 = pixel-major, 2 = weights resident in
                                   // TMEM (conv_tct.cu; 32-pixel-wide images, else pixel-major)
       TUNE_TC_PAIR = 4,           // pixel-major conv on a CTA pair (cta_group::2, M = 256): 0 off, 1 on, 2 for C >= 128 (default)
       TUNE_WAIT_BACKOFF = 5,      // nanosleep back-off (ns, first step) of waiting epilogue / producer warps; 0 = tight poll
       TUNE_PDL = 6,               // programmatic dependent launch of the tcgen05 kernels (prologue overlaps the predecessor's tail)
       TUNE_WGRAD_MULTICAST = 7,   // weight-gradient GEMM: cluster of the tap groups, gout box multicast (one L2 read per cluster)
       TUNE_TCT_BAND = 8,          // TMEM-resident-weight conv: image rows per work item (0 = 16 / 8 / 4 by image height)
       TUNE_TCT_DEBUG = 9,         // TMEM-resident-weight conv: decomposition switches (timing experiments only; results are garbage)
       TUNE_TCT_PRODUCTS = 10,     // TMEM-resident-weight conv: hi/lo products formed, 3 (default) or 4
       TUNE_MMA_WARP_HIGH = 11,    // pixel-major convs: TMA / MMA roles on the highest physical warps (scheduler priority)
       TUNE_WGRAD64_PRODUCTS = 12, // C = 64 weight gradient: 4 hi/lo products (default) or 3 (roles swapped; measured slower)
       TUNE_MNIST_FUSED = 13,      // MNIST right-hand side forward: whole solve in one persistent tcgen05 launch (1, default) or the SIMT multi-launch path
       TUNE_UNIFORM_ISSUE = 14,    // warp-uniform MMA issue loop instead of the one-lane loop of round 1: bit 0 = CTA-pair conv (default on: -2..3 %),
                                   // bit 1 = weight gradient (default off: no gain measured)
       TUNE_GN_BLOCK = 15,         // GroupNorm of large states (CIFAR GN / LN / IN right-hand sides): one CTA per sample (1, default)
                                   // or the warp-per-(sample, group) kernels written for the MNIST state (0)
       TUNE_TCP2_HALF_STAGE = 16,  // C = 128 CTA-pair conv: half-size epilogue stage (two passes) + a third activation stage (halo form:
                                   // + two more weight stages); default 0 since the halo form removed the activation waits it was for
       TUNE_TCP2_HALO = 17,        // C = 128, 16-pixel-wide images: one staged halo tile per c_in chunk serves all nine taps (1, default)
       TUNE_WGRAD_HTAPS = 18,      // weight gradient: a CTA owns a vertical tap, the three horizontal taps are N atoms 128 B apart
                                   // in one staged copy (1, default: -17 % L2 -> SM bytes, C = 64 launch -10 % under the power cap) or
                                   // a CTA owns a horizontal tap with its own shifted box (0, round 1)
       TUNE_PEER_FORM = 19,        // gradient exchange over peer memory (peer.cu): 0 = two-shot from three ranks on (default), 1 = always
                                   // one-shot (every rank reads every gradient), 2 = always two-shot (reduce-scatter + peer stores)
       TUNE_COUNT };
int tune_get(int which);

struct ConvShape { int B, H, W, C; };

// Launch `kern` with the programmatic-stream-serialization attribute when the "pdl" option is on (the kernel must call
// ptx::pdl_wait() before it touches anything its predecessor wrote), plainly otherwise.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tune_get(TUNE_PDL) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- elementwise.cu ----
void launch_act_split(const float* x, const float* mul, int act, float scale, __nv_bfloat16* split, float* dact,
                      int B, int H, int W, int C, cudaStream_t st);
void launch_act_split_sliced(const float* x, const float* mul, int act, const float* scales_host, int n_slices,
                             __nv_bfloat16* split, float* dact, int B, int H, int W, int C, cudaStream_t st);
void launch_pack_w_simt(const float* w, float* out, int C, int Cin_total, int skip_in, int transpose, cudaStream_t st);
void launch_pack_w_tc(const float* w, __nv_bfloat16* out, int C, int transpose, cudaStream_t st);
void launch_wgrad_reduce(const float* partial, int nparts, float* grad_w, int C, int accumulate, cudaStream_t st,
                         int cin_total = 0, int skip_in = 0);

// ---- train_aux.cu: deterministic dot product (*out += scale * <a, b>) ----
size_t dot_scratch_bytes();
void launch_dot_accumulate(const float* a, const float* b, size_t n, double scale, double* out, double* scratch, cudaStream_t st);
void launch_dot_bcast_accumulate(const float* a, const float* b, size_t n, size_t period, double scale, double* out,
                                 double* scratch, cudaStream_t st);

// ---- groupnorm.cu (MNIST right-hand side) ----
int launch_groupnorm_epi(const float* x, const float* gamma, const float* beta, const EpiParams& epi, ConvShape s,
                         int groups, float eps, cudaStream_t st);
int launch_groupnorm_bwd_epi(const float* x, const float* gamma, const float* beta, const float* dy, float dy_scale, int relu,
                             const EpiParams& epi, float* dgamma_part, float* dbeta_part, int accumulate, ConvShape s,
                             int groups, float eps, cudaStream_t st);
void launch_sum_over_batch(const float* part, float* out, int B, int C, cudaStream_t st);
void launch_concat_aux_grad(const float* dP, float t, float* gw, float* gb, int accumulate, ConvShape s, cudaStream_t st);
void launch_time_tapmap(const float* w, float* tapmap, int H, int W, int C, cudaStream_t st);

// ---- netlayers.cu (stem, strided residual block re-indexing) ----
int launch_stem_fwd(const float* x, const float* w, int act, float* y, float* dact, int B, int H, int W, int C,
                    cudaStream_t st);
int stem_wgrad_blocks();
int launch_stem_wgrad(const float* gy, const float* dact, const float* x, float* partial, float* gw, int B, int H, int W,
                      int C, cudaStream_t st);
int launch_stem_dgrad(const float* gy, const float* dact, const float* w, float* gx, int B, int H, int W, int C,
                      cudaStream_t st);
void launch_s2d_act_split(const float* x, int act, __nv_bfloat16* T0, __nv_bfloat16* T1, __nv_bfloat16* Tsc, float* G0,
                          int B, int H, int W, int Ci, cudaStream_t st);
void launch_d2s_grad(const float* gT0, const float* gT1, const float* gTsc, const float* G0, float* gx, int B, int H, int W,
                     int Ci, cudaStream_t st);
void launch_down_weights_build(const float* w1, const float* wsc, float* Wd, int Ci, int Co, cudaStream_t st);
void launch_down_weights_gather(const float* gWd, float* gw1, float* gwsc, int Ci, int Co, cudaStream_t st);

// ---- conv_simt.cu : plain fp32 FFMA engine (any C multiple of 4, any H, W) ----
//   out-epilogue(conv3x3(split_in, w_packed[tap][ci][co]))
int launch_conv3x3_simt(const __nv_bfloat16* split_in, const float* w_packed, const EpiParams& epi,
                        ConvShape s, cudaStream_t st);
//   partial[nparts][tap][ci][co] = per-slice sums of in(shifted)[ci] * gout[co]; returns nparts via *nparts_out
int launch_wgrad3x3_simt(const __nv_bfloat16* split_gout, const __nv_bfloat16* split_in, float* partial,
                         int* nparts_out, ConvShape s, cudaStream_t st);
int wgrad_simt_nparts(ConvShape s);

// ---- conv_tc.cu : tcgen05 / TMEM / TMA implicit-GEMM engine ----
bool tc_shape_supported(int C, int H, int W);
size_t tc_packed_weight_bytes(int C);
int launch_conv3x3_tc(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi,
                      ConvShape s, cudaStream_t st);

// ---- conv_tcp.cu : the same engine in pixel-major form (pixels on M, 3 hi/lo products, vector epilogue) ----
size_t tcp_packed_weight_bytes(int C);
void launch_pack_w_tcp(const float* w, __nv_bfloat16* out, int C, int transpose, cudaStream_t st);
int launch_conv3x3_tcp(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi,
                       ConvShape s, cudaStream_t st);
// the same on a CTA pair (conv_tcp2.cu): tcgen05.mma.cta_group::2, weights shared by the pair; same packed weights
bool tcp2_shape_supported(int B, int C, int H, int W);
int launch_conv3x3_tcp2(const __nv_bfloat16* split_in, const __nv_bfloat16* w_tiles, const EpiParams& epi,
                        ConvShape s, cudaStream_t st);
// which form MSB_ENGINE_TCGEN05 runs for a shape: 0 = channel-major (conv_tc.cu), 1 = pixel-major (conv_tcp.cu / conv_tcp2.cu),
// 2 = weights resident in TMEM (conv_tct.cu)   (odeblock.cu; env MSB_TC_CONV=cm|pm forces one of the first two)
int tc_form(int C, int H, int W);

// ---- conv_tct.cu : C = 64, weights resident in tensor memory as the A operand, activations on N from a row ring ----
bool tct_shape_supported(int C, int H, int W);
size_t tct_packed_weight_bytes();
void launch_pack_w_tct(const float* w, void* out, int transpose, cudaStream_t st);
int launch_conv3x3_tct(const __nv_bfloat16* split_in, const void* wpacked, const EpiParams& epi, ConvShape s, cudaStream_t st);
int tct_products();

// ---- wgrad_tc.cu ----
bool wgrad_tc_supported(ConvShape s);
int wgrad_tc_nparts(ConvShape s);
// accumulate != 0: partial[slot] += result (each CTA owns its slots: deterministic)
int launch_wgrad3x3_tc(const __nv_bfloat16* split_gout, const __nv_bfloat16* split_in, float* partial,
                       int* nparts_out, int accumulate, ConvShape s, cudaStream_t st);

}  // namespace msb
