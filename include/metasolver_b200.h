/*
 * metasolver_b200 -- C ABI of the B200-native meta-solver ODE-block hot path.
 *
 * The reference (juliagusak/neural-ode-metasolver) is pure Python/PyTorch and has no FFI layer;
 * the seam it offers is the `sopa` Python API.  This header is the C boundary that API binds to
 * (ctypes stub: neural-ode-metasolver_b200/_cabi.py).  Plain pointers and
 * sizes only -- no torch types.  Every entry point:
 *   - borrows all device buffers for the duration of the call (caller allocates / frees),
 *   - enqueues its kernels on the caller's CUDA stream and never synchronises the host,
 *   - returns 0 on success, non-zero on error (text via msb_last_error(), thread-local),
 *   - refuses unsupported (rhs kind, activation, shape) combinations: there is no cuDNN / CPU
 *     fallback behind this ABI.
 *
 * Device layouts
 *   state tensors : fp32  NHWC            [B][H][W][C]              ("channels-last")
 *   split tensors : bf16  [B][H][2][W][C] plane 0 = hi = bf16(x), plane 1 = lo = bf16(x - hi)
 *                   (the form in which every convolution operand lives in HBM; hi+lo carries
 *                    ~16 significand bits, the tensor cores multiply all four hi/lo products
 *                    and accumulate in fp32)
 *   conv weights  : fp32  OIHW            [C_out][C_in][3][3]       (PyTorch's own layout)
 *
 * What each function replaces in the reference (paths relative to the reference root):
 *   msb_odeblock_forward   RKParametricSolver.integrate            sopa/src/solvers/rk_parametric.py:89-113
 *                          + RK*._make_step                        sopa/src/solvers/rk_parametric_order2stage2.py:87-93,
 *                                                                  ..order3stage3.py:96-103, ..order4stage4.py:184-192, euler.py:63-68
 *                          + PreBasicBlock2 / BasicBlock2.forward  sopa/src/models/odenet_cifar10/layers.py:148-161, 108-121
 *                          + ODEfunc / ConcatConv2d.forward        sopa/src/models/odenet_mnist/layers.py:158-171, 250-253
 *   msb_odeblock_backward  torch.autograd through all of the above (discretize-then-optimize;
 *                          consumers: examples/cifar10/train_and_attack.py:310-311,
 *                          MegaAdversarial/src/attacks/fgsm.py:34-36,98, pgd.py:44-46)
 */
#ifndef METASOLVER_B200_H_
#define METASOLVER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSB_ABI_VERSION 6   /* v3: + tuning options, attack / SGD steps, tableau gradients, MSB_RHS_PREACT_GN; v4: + network head (pool + FC, cross-entropy); v5: + MSB_RHS_POSTACT_GN, stacked tableau gradients, msb_augment_batch (structs unchanged); v6: + msb_peer_* (gradient all-reduce + SGD update as one kernel over peer memory) */
#define MSB_MAX_STAGES 4

/* right-hand-side families */
enum { MSB_RHS_PREACT_NF = 0,   /* conv2(act(conv1(act(x))))          cifar10/layers.py:148-161 */
       MSB_RHS_POSTACT_NF = 1,  /* act(conv2(act(conv1(x))))          cifar10/layers.py:108-121 */
       MSB_RHS_MNIST_GN_T = 2,  /* GN-ReLU-cconv(t)-GN-ReLU-cconv(t)-GN  mnist/layers.py:158-171 */
       MSB_RHS_PREACT_GN = 3,   /* conv2(act(GN2(conv1(act(GN1(x))))))  cifar10/layers.py:148-161 with the
                                   'GN' / 'LN' / 'IN' normalisations of cifar10/utils.py:26-36 (all nn.GroupNorm).
                                   Parameters travel in MsbMnistParams: norm_w/b[0..1], conv_w[0..1] = OIHW (C,C,3,3)
                                   without bias (conv_b, norm_w/b[2] ignored). */
       MSB_RHS_POSTACT_GN = 4   /* act(GN2(conv2(act(GN1(conv1(x))))))  BasicBlock2 (cifar10/layers.py:108-121) with the same
                                   per-sample normalisations; parameters as for MSB_RHS_PREACT_GN. */ };
/* activations */
enum { MSB_ACT_NONE = 0, MSB_ACT_GELU_ERF = 1, MSB_ACT_RELU = 2 };
/* GEMM engines */
enum { MSB_ENGINE_AUTO = 0,     /* tcgen05 when the shape is covered by it, else the SIMT CUDA engine */
       MSB_ENGINE_TCGEN05 = 1,  /* implicit-GEMM on tcgen05/TMEM fed by TMA (sm_100a) */
       MSB_ENGINE_SIMT = 2      /* plain fp32 FFMA CUDA kernels (any shape; validation / small shapes) */ };

/* Butcher tableau of one solver of a stacked solver axis (same meaning as MsbOdeDesc.c/b/w). */
typedef struct MsbTableau {
    float c[MSB_MAX_STAGES];
    float b[MSB_MAX_STAGES];
    float w[MSB_MAX_STAGES * MSB_MAX_STAGES];
} MsbTableau;
#define MSB_MAX_SOLVERS 8

/* One ODE-block integration problem.  All scalar arrays are HOST values, read during the call. */
typedef struct MsbOdeDesc {
    int32_t rhs_kind;                 /* MSB_RHS_* */
    int32_t act;                      /* MSB_ACT_* used inside the RHS */
    int32_t engine;                   /* MSB_ENGINE_* */
    int32_t batch, height, width, channels;
    int32_t n_steps;                  /* number of RK steps = len(grid) - 1 */
    int32_t stages;                   /* 1..4 */
    float c[MSB_MAX_STAGES];          /* Butcher nodes   (rk_parametric.py:68-75) */
    float b[MSB_MAX_STAGES];          /* Butcher weights */
    float w[MSB_MAX_STAGES * MSB_MAX_STAGES]; /* row-major lower-triangular stage matrix w[i][j] */
    const float* time_grid;           /* HOST pointer, n_steps+1 grid points (rk_parametric.py:93-96) */
    int32_t save_tape;                /* 1: record what backward needs into `tape` */
    /* Stacked solver axis (solver ensembling, cifar10/layers.py:190-205; model ensembling, fgsm.py:135-143):
     * n_solvers = K > 1 integrates K equal slices of the batch, slice s (images [s*batch/K, (s+1)*batch/K))
     * with tableau solver_tableaus[s], in the SAME launches -- all solvers share stages, n_steps and time_grid.
     * 0 or 1: one solver, tableau = c/b/w above. */
    int32_t n_solvers;
    const MsbTableau* solver_tableaus; /* HOST pointer, n_solvers entries; ignored when n_solvers <= 1 */
} MsbOdeDesc;

/* Extra parameters of the MNIST RHS (device pointers, fp32). */
typedef struct MsbMnistParams {
    const float* norm_w[3];           /* GroupNorm gamma [C]            mnist/layers.py:152-157 */
    const float* norm_b[3];           /* GroupNorm beta  [C] */
    const float* conv_w[2];           /* [C][C+1][3][3] (input channel 0 is the time channel) */
    const float* conv_b[2];           /* [C] */
    int32_t groups;                   /* min(32, C)                      mnist/layers.py:208-209 */
    float eps;                        /* 1e-5 */
} MsbMnistParams;

/* Gradients of the MNIST right-hand-side parameters (device pointers, fp32, shapes of MsbMnistParams; overwritten). */
typedef struct MsbMnistGrads {
    float* norm_w[3];
    float* norm_b[3];
    float* conv_w[2];                 /* [C][C+1][3][3], incl. the time channel */
    float* conv_b[2];
} MsbMnistGrads;

int         msb_abi_version(void);
const char* msb_last_error(void);
/* sizeof() of the ABI structs as this library was compiled (0 MsbOdeDesc, 1 MsbTableau, 2 MsbMnistParams,
 * 3 MsbMnistGrads, 4 MsbDownDesc): lets a foreign-language binding verify its struct layout at load time. */
size_t      msb_sizeof(int which);
/* 1 if `device` can run the tcgen05 engine (compute capability 10.x), 0 if not, <0 on error. */
int         msb_device_supports_tcgen05(int device);
/* 1 if (C,H,W) is covered by the tcgen05 engine. */
int         msb_shape_supports_tcgen05(int channels, int height, int width);

/* Scratch (not preserved between calls) and tape (forward -> backward) sizes in bytes. */
size_t msb_odeblock_workspace_bytes(const MsbOdeDesc* d);
size_t msb_odeblock_tape_bytes(const MsbOdeDesc* d);
size_t msb_odeblock_bwd_workspace_bytes(const MsbOdeDesc* d);

/* y_out = state at time_grid[n_steps] starting from x at time_grid[0].  x, y_out: fp32 NHWC.
 * w1, w2: fp32 OIHW conv weights of the RHS (CIFAR families).  `mnist` non-NULL only for
 * MSB_RHS_MNIST_GN_T (then w1/w2 are ignored).  `tape` may be NULL when save_tape == 0. */
int msb_odeblock_forward(const MsbOdeDesc* d, const float* x, const float* w1, const float* w2,
                         const MsbMnistParams* mnist, float* y_out,
                         void* workspace, size_t workspace_bytes, void* tape, size_t tape_bytes,
                         void* cuda_stream);

/* Discretize-then-optimize gradient of the same integration.  grad_y, grad_x: fp32 NHWC.
 * grad_w1/grad_w2 (fp32 OIHW) are OVERWRITTEN when non-NULL; pass NULL for the input-gradient
 * only mode used by FGSM/PGD (pgd.py:44-46). */
int msb_odeblock_backward(const MsbOdeDesc* d, const float* grad_y, const float* w1, const float* w2,
                          const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1, float* grad_w2,
                          void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Same for MSB_RHS_MNIST_GN_T (tape recorded by msb_odeblock_forward with save_tape = 1).  `grads` NULL = input
 * gradient only.  Replaces torch.autograd through ODEfunc / ConcatConv2d / GroupNorm, mnist/layers.py:158-171, 250-253. */
/* The same backward passes, additionally ACCUMULATING the gradient w.r.t. the Butcher coefficients into
 * grad_tableau (device, MSB_TABLEAU_GRAD_DOUBLES doubles: dL/db_i [M], dL/dw_ij row-major [M*M], dL/dc_i [M],
 * M = MSB_MAX_STAGES) -- what autograd yields for solver.u / solver.v after unfreeze_params()
 * (rk_parametric_order2stage2.py:104-109) once chained through the closed-form tableau on the host.  dL/dc_i is
 * non-zero only for the time-dependent MNIST right-hand side (t_i = t_n + c_i dt, order2stage2.py:81-86).  With a stacked
 * solver axis (n_solvers = K > 1; msb_odeblock_backward_tableau only) grad_tableau holds K consecutive blocks of
 * MSB_TABLEAU_GRAD_DOUBLES doubles, block s reduced over the images of slice s.  Needs
 * msb_odeblock_bwd_workspace_bytes_tableau() bytes of workspace. */
#define MSB_TABLEAU_GRAD_DOUBLES (MSB_MAX_STAGES + MSB_MAX_STAGES * MSB_MAX_STAGES + MSB_MAX_STAGES)
size_t msb_odeblock_bwd_workspace_bytes_tableau(const MsbOdeDesc* d);
int msb_odeblock_backward_tableau(const MsbOdeDesc* d, const float* grad_y, const float* w1, const float* w2,
                                  const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1, float* grad_w2,
                                  double* grad_tableau, void* workspace, size_t workspace_bytes, void* cuda_stream);
int msb_odeblock_backward_mnist_tableau(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mnist,
                                        const void* tape, size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads,
                                        double* grad_tableau, void* workspace, size_t workspace_bytes, void* cuda_stream);

int msb_odeblock_backward_mnist(const MsbOdeDesc* d, const float* grad_y, const MsbMnistParams* mnist, const void* tape,
                                size_t tape_bytes, float* grad_x, const MsbMnistGrads* grads, void* workspace,
                                size_t workspace_bytes, void* cuda_stream);

/* ---- non-ODE layers of the CIFAR networks on the same engines ------------------------
 * Stem: y = act(conv3x3(x, w)), 3 input channels -> `channels`, stride 1, pad 1, no bias
 * (MetaNODE.forward, sopa/src/models/odenet_cifar10/layers.py:411-413).  x: fp32 NHWC (B,H,W,3);
 * y, dact_out (= act'(pre-activation), NULL when no backward will follow): fp32 NHWC (B,H,W,channels). */
int msb_stem_forward(const float* x_nhwc, const float* w_oihw, int act, float* y, float* dact_out,
                     int batch, int height, int width, int channels, void* cuda_stream);
size_t msb_stem_backward_workspace_bytes(int channels);
/* grad_w (OIHW, overwritten) may be NULL (input-gradient-only attacks); grad_x (NHWC, 3 channels) may be NULL. */
int msb_stem_backward(const float* grad_y, const float* dact, const float* x_nhwc, const float* w_oihw,
                      float* grad_w, float* grad_x, int batch, int height, int width, int channels,
                      void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Strided pre-activation residual block  y = conv2(act(conv1_s2(act(x)))) + conv1x1_s2(x)
 * (PreBasicBlock with stride 2 and the 1x1 shortcut, layers.py:54-81, NF norm): in_channels -> out_channels
 * = 2*in_channels, (H,W) -> (H/2,W/2).  w1: (Co,Ci,3,3), w2: (Co,Co,3,3), wsc: (Co,Ci,1,1), fp32 OIHW. */
typedef struct MsbDownDesc {
    int32_t act;                      /* MSB_ACT_* */
    int32_t engine;                   /* MSB_ENGINE_* */
    int32_t batch, height, width;     /* input size (even height and width) */
    int32_t in_channels, out_channels;
    int32_t save_tape;
} MsbDownDesc;
size_t msb_downblock_workspace_bytes(const MsbDownDesc* d);
size_t msb_downblock_tape_bytes(const MsbDownDesc* d);
size_t msb_downblock_bwd_workspace_bytes(const MsbDownDesc* d);
int msb_downblock_forward(const MsbDownDesc* d, const float* x, const float* w1, const float* w2, const float* wsc,
                          float* y_out, void* workspace, size_t workspace_bytes, void* tape, size_t tape_bytes,
                          void* cuda_stream);
/* grad_w1 / grad_w2 / grad_wsc: all given (overwritten) or all NULL (input-gradient only). */
int msb_downblock_backward(const MsbDownDesc* d, const float* grad_y, const float* w1, const float* w2,
                           const float* wsc, const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1,
                           float* grad_w2, float* grad_wsc, void* workspace, size_t workspace_bytes,
                           void* cuda_stream);

/* ---- building blocks, exported so the parity tests can exercise each kernel through the C ABI ---- */

/* split[B][H][2][W][C] = hi/lo(act(x)) ; dact (optional, fp32 NHWC) = act'(x). */
int msb_act_split(const float* x, int act, void* split_out, float* dact_out,
                  int batch, int height, int width, int channels, void* cuda_stream);

/* out = conv3x3(split_in, w) (stride 1, pad 1, no bias), fp32 NHWC.  transpose != 0 computes the
 * input-gradient convolution (weights transposed and rotated by 180 degrees). */
int msb_conv3x3(const void* split_in, const float* w_oihw, float* out, int transpose, int engine,
                int batch, int height, int width, int channels,
                void* workspace, size_t workspace_bytes, void* cuda_stream);
size_t msb_conv3x3_workspace_bytes(int channels);

/* grad_w[O][I][3][3] = sum over pixels of grad_out (x) shifted input; both operands split tensors. */
int msb_wgrad3x3(const void* split_grad_out, const void* split_in, float* grad_w_oihw, int engine,
                 int batch, int height, int width, int channels,
                 void* workspace, size_t workspace_bytes, void* cuda_stream);
size_t msb_wgrad3x3_workspace_bytes(int channels, int engine);

/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
uint64_t msb_launch_count(void);

/* ---- callers either side of the path -------------------------------------------
 * Elementwise steps of the adversarial attacks as ONE kernel each, bit-identical to the reference's torch
 * calls (MegaAdversarial/src/attacks/fgsm.py:27-40,93-105, pgd.py:28-53).  All tensors are fp32 images of
 * n_elements = B*channels*hw values, NCHW-contiguous (channels_last = 0) or channels-last memory (1).
 * chan_consts = 4 x channels floats [mean|lower][std|upper][eps][alpha] (per channel), may be NULL where unused.
 *   UNNORMALIZE / NORMALIZE : out = (a - c0) / c1
 *   FGSM_STEP   : out = clamp01(a + eps*sign(grad));                       normalize_out: (out - c0)/c1
 *   PGD_STEP    : out = clamp01(clamp(a + step*sign(grad), ref-eps, ref+eps));   normalize_out as above
 *   FGSMR_INIT  : out = clamp(c2 - (2 c2)*a, c0 - ref, c1 - ref)           (a = U[0,1) noise, ref = x)
 *   FGSMR_STEP  : d = clamp(clamp(a + c3*sign(grad), -c2, c2), c0 - ref, c1 - ref);  out = normalize_out ? ref + d : d
 */
enum { MSB_ATTACK_UNNORMALIZE = 0, MSB_ATTACK_NORMALIZE = 1, MSB_ATTACK_FGSM_STEP = 2, MSB_ATTACK_PGD_STEP = 3,
       MSB_ATTACK_FGSMR_INIT = 4, MSB_ATTACK_FGSMR_STEP = 5 };
#define MSB_ATTACK_MAX_CHANNELS 4
int msb_attack_step(int kind, const float* a, const float* grad, const float* ref, float* out, int64_t n_elements,
                    int channels, int hw, int channels_last, float eps, float step, int normalize_out,
                    const float* chan_consts /* host */, void* cuda_stream);

/* Training-input transform on the device (replaces the per-sample torchvision pipeline RandomCrop(32, padding=4) +
 * RandomHorizontalFlip + ToTensor + Normalize of sopa/src/models/odenet_cifar10/data.py:40-46, and the test-time
 * ToTensor + Normalize of :54-57 when dx = dy = flip = NULL):
 *   out[b][y][x][c] = (pad0(images_u8[index[b]])[y + dy[b]][xf + dx[b]][c] / 255 - mean[c]) / std[c],
 *   xf = flip[b] ? width-1-x : x,   pad0 = the uint8 image zero-padded by `padding` on every side,  dx, dy in 0..2*padding.
 * images_u8: uint8 [N][height][width][channels] (the dataset, resident in HBM); index (int64, NULL = 0..batch-1), dx, dy
 * (int32, NULL = padding: centred window), flip (uint8, NULL = none): device arrays of `batch` entries (the random draws
 * stay with the caller's generator); mean, std: HOST arrays of `channels` floats.  out: fp32 NHWC = the stem's input
 * layout.  Same fp32 operations in the same order as torchvision: bit-identical. */
int msb_augment_batch(const uint8_t* images_u8, const int64_t* index, const int32_t* dx, const int32_t* dy, const uint8_t* flip,
                      int batch, int height, int width, int channels, int padding, const float* mean /* host */,
                      const float* std /* host */, float* out_nhwc, void* cuda_stream);

/* SGD with momentum and weight decay (torch.optim.SGD semantics, no dampening / nesterov;
 * examples/cifar10/train_and_attack.py:98-99,322) over ONE flat fp32 buffer:
 *   g = grads*grad_scale + weight_decay*p;  buf = first_step ? g : momentum*buf + g;  p -= lr*buf
 * grad_scale carries the 1/world average of the data-parallel all-reduce.  momentum_buf may be NULL if momentum == 0. */
int msb_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum,
                 float weight_decay, float grad_scale, int first_step, void* cuda_stream);

/* Data-parallel exchange step over PEER MEMORY (NVLink 5 / NVSwitch), SURVEY 8(e) + 8(f-4): the ONE gradient all-reduce of
 * a training step (reference: single-process; north_star: "a single NCCL-over-NVLink gradient allreduce per step") fused
 * with the optimizer update that consumes it (examples/cifar10/train_and_attack.py:98-99,322) -- one launch per rank, no NCCL
 * call, no intermediate buffer.
 *   msb_peer_alloc   cudaMalloc of [MSB_PEER_HEADER_BYTES header | payload], zero-filled, + its CUDA IPC handle
 *                    (MSB_PEER_HANDLE_BYTES bytes the caller ships to the other ranks of the node, e.g. all_gather_object)
 *   msb_peer_open    map another rank's buffer (enables peer access between the two devices); msb_peer_close unmaps it
 *   msb_peer_free    release the own buffer (after every peer has closed it)
 *   msb_peer_status  synchronous read of the own header: error word (0 = ok, MSB_PEER_ERR_*) and the last finished epoch
 *   msb_peer_allreduce_sgd
 *       bases[world]: exchange buffers in rank order, bases[rank] = the own one.  The payload of every buffer holds the rank's
 *       flat fp32 gradient at float 0 and, if result_offset >= 0, a result array of the same indexing at float result_offset.
 *       This launch covers floats [offset, offset + n).  Per element, on every rank:
 *           s = (((g_0 + g_1) + g_2) + ... + g_{world-1}) * grad_scale            fixed order: bitwise identical on all ranks
 *           result[i] = s                                       (when write_result, and always in the two-shot form)
 *           g = s + weight_decay*p;  buf = first_step ? g : momentum*buf + g;  p -= lr*buf     (if params != NULL; = msb_sgd_step)
 *       params / momentum_buf point at element `offset` of their flat arrays (own memory).  Two forms of the same launch:
 *       one-shot (every rank reads all gradients) and, with a result array and from three ranks on, two-shot (rank r reduces
 *       slice r and stores the average into every rank's result array, then updates from its own result array); option
 *       "peer_form" forces one.  All ranks must issue the same sequence of launches.  The launch starts with a system-scope
 *       handshake (all gradients complete) and contains a second one (nobody reads a gradient any more / all slices have
 *       landed): when it has finished, the stream's successor may overwrite the gradient, and the result array is valid until
 *       the next exchange launch of this rank.  The epoch ordering the messages is kept in the header, so the call can be
 *       captured in a CUDA graph.  Waits are bounded by timeout_ms (0 = 10 s): on expiry the header's error word is set and
 *       the launch completes with undefined results instead of hanging. */
#define MSB_PEER_MAX_RANKS 16
#define MSB_PEER_HANDLE_BYTES 64
#define MSB_PEER_HEADER_BYTES 1024
enum { MSB_PEER_OK = 0, MSB_PEER_ERR_READY_TIMEOUT = 1, MSB_PEER_ERR_DONE_TIMEOUT = 2 };
int msb_peer_alloc(size_t payload_bytes, void** base, unsigned char* handle);
int msb_peer_open(const unsigned char* handle, void** base);
int msb_peer_close(void* base);
int msb_peer_free(void* base);
int msb_peer_status(const void* own_base, unsigned* error_word, unsigned* epoch);
int msb_peer_allreduce_sgd(void* const* bases, int world, int rank, int64_t offset, int64_t n, int64_t result_offset,
                           int write_result, float* params, float* momentum_buf, float lr, float momentum, float weight_decay,
                           float grad_scale, int first_step, unsigned timeout_ms, void* cuda_stream);

/* Network head of MetaNODE: AdaptiveAvgPool2d((1,1)) + Flatten + Linear (sopa/src/models/odenet_cifar10/layers.py:390-392,425)
 * on an NHWC fp32 map, and its gradient.  pooled[batch][channels] is an output of the forward (saved for the backward).
 *   logits[n][k] = bias[k] + sum_c W[k][c] * mean_p x[n][p][c]
 * backward: dx[n][p][c] = (sum_k dlogits[n][k] W[k][c]) / hw   (dx may be NULL),
 *           dw[k][c] = sum_n dlogits[n][k] pooled[n][c], dbias[k] = sum_n dlogits[n][k]   (dw NULL = no parameter gradients;
 *           fixed summation order: bitwise reproducible).  channels must be a multiple of 4. */
int msb_pool_fc_forward(const float* x_nhwc, const float* w, const float* bias, float* pooled, float* logits, int batch,
                        int hw, int channels, int classes, void* cuda_stream);
int msb_pool_fc_backward(const float* dlogits, const float* w, const float* pooled, float* dx_nhwc, float* dw, float* dbias,
                         int batch, int hw, int channels, int classes, void* cuda_stream);
/* Mean cross-entropy over the batch (nn.CrossEntropyLoss / F.cross_entropy with default reduction;
 * examples/cifar10/train_and_attack.py:303-311, fgsm.py:33, pgd.py:43): loss[0] = mean_n (logsumexp(z_n) - z_n[y_n]);
 * lse[batch] is saved for the backward: dlogits = (softmax(z) - onehot(y)) * grad_loss[0] / batch.  labels are int64. */
int msb_cross_entropy_forward(const float* logits, const int64_t* labels, float* loss, float* lse, int batch, int classes,
                              void* cuda_stream);
int msb_cross_entropy_backward(const float* logits, const int64_t* labels, const float* lse, const float* grad_loss,
                               float* dlogits, int batch, int classes, void* cuda_stream);

/* Tuning options (process-wide; defaults are the measured best).  Names:
 *   "epi_l2_prefetch"  distance, in tiles, at which the tcgen05 convolutions bulk-prefetch their epilogue
 *                      operands (y, k_j, act') into L2; 0 = off          (env MSB_EPI_L2_PREFETCH)
 *   "tc_resident"      1 = channel-major C=64 convolution keeps its weights resident in shared memory
 *                                                                        (env MSB_TC_RESIDENT)
 *   "tcp_epi_warps"    epilogue warps of the pixel-major convolution, 8 or 16      (env MSB_TCP_EPI_WARPS)
 *   "tc_form_c64"      tcgen05 convolution form for 64 channels: 0 = channel-major (4 hi/lo products),
 *                      1 = pixel-major (3 products), 2 (default) = weights resident in tensor memory as the A operand,
 *                      activations from a ring of image rows (32-pixel-wide images; pixel-major elsewhere)
 *                                                                        (env MSB_TC_FORM_C64)
 *   "tct_band"         image rows per work item of form 2 (0 = 16 / 8 / 4 by image height)   (env MSB_TCT_BAND)
 *   "tct_products"     hi/lo products of form 2: 4 (default) or 3 (an M = 64 MMA for the lo plane)  (env MSB_TCT_PRODUCTS)
 *   "tct_debug"        decomposition switches of form 2 for timing experiments (results are garbage; 0 = off)
 *   "mma_warp_high"    1 = pixel-major convolutions place their TMA / MMA warps on the highest warp ids (env MSB_MMA_WARP_HIGH)
 *   "tc_pair"          pixel-major convolution on CTA pairs (tcgen05.mma.cta_group::2, M = 256, weights shared by
 *                      the pair): 0 = off, 1 = every shape with an even tile count, 2 = only C >= 128 (default)
 *                                                                        (env MSB_TC_PAIR)
 *   "wait_backoff_ns"  first nanosleep step of waiting epilogue / TMA-producer warps (doubles up to 8x); 0 = tight
 *                      mbarrier polling                                  (env MSB_WAIT_BACKOFF_NS)
 *   "pdl"              1 (default) = the tcgen05 kernels are launched with programmatic stream serialization: their
 *                      prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the predecessor's tail,
 *                      griddepcontrol.wait orders every access to the predecessor's outputs   (env MSB_PDL)
 *   "wgrad_multicast"  1 = weight-gradient GEMM as clusters of the tap groups with the shared gout box loaded once by
 *                      TMA multicast (measured slower with the current two-stage ring; default 0)  (env MSB_WGRAD_MULTICAST)
 *   "wgrad64_products" hi/lo products of the C = 64 weight gradient: 4 (default) or 3 (roles swapped; measured slower)
 *   "mnist_fused"      1 (default) = the MNIST ODE-block forward runs as ONE persistent tcgen05 launch; 0 = multi-launch path
 *   "uniform_issue"    warp-uniform MMA issue loops: bit 0 = CTA-pair convolution (default on), bit 1 = weight gradient
 *   "gn_block"         1 (default) = GroupNorm of large states (CIFAR 'GN' / 'LN' / 'IN' right-hand sides) as one CTA per
 *                      sample; 0 = the warp-per-(sample, group) kernels of the MNIST state      (env MSB_GN_BLOCK)
 *   "wgrad_htaps"      1 (default) = weight gradient with one vertical tap per CTA and the three horizontal taps as N atoms
 *                      128 B apart in ONE staged copy of the input rows; 0 = one horizontal tap per CTA with its own shifted
 *                      copy (round 1)                                     (env MSB_WGRAD_HTAPS)
 *   "tcp2_halo"        1 (default) = C = 128 convolution on 16-pixel-wide images stages ONE halo tile per c_in chunk for all
 *                      nine taps (horizontal taps = the same tile read 128 B further); 0 = three shifted copies (env MSB_TCP2_HALO)
 *   "peer_form"        gradient exchange over peer memory (msb_peer_allreduce_sgd): 0 (default) = two-shot from three ranks on,
 *                      1 = always one-shot, 2 = always two-shot           (env MSB_PEER_FORM)
 *   "tcp2_half_stage"  1 = half-size epilogue stage of the C = 128 pair convolution, the memory going to deeper operand rings
 *                      (default 0)                                        (env MSB_TCP2_HALF_STAGE)
 * Results do not depend on any option except the products formed (tc_form_c64, tct_products, wgrad64_products: last-bit
 * differences), the summation orders of mnist_fused / gn_block (same), and tct_debug.  Returns 0, or -1 for an unknown name. */
int msb_set_option(const char* name, int value);
int msb_get_option(const char* name, int* value);

/* Optional per-launch timing for the roofline report: when enabled, CUDA events are recorded on the
 * caller's stream around every convolution-engine launch.  msb_profile_read() waits for the recorded
 * events and returns the summed duration, the summed algorithmic flops (2*M*N*K of the convolution)
 * and the number of launches of one kind.  msb_profile_enable() also clears the records. */
enum { MSB_PROF_CONV = 0,   /* forward / input-gradient convolutions */
       MSB_PROF_WGRAD = 1   /* weight-gradient GEMMs */ };
int msb_profile_enable(int on);
int msb_profile_read(int kind, double* total_ms, double* total_flops, int64_t* count);
/* Tensor-core flops actually EXECUTED by the recorded launches of `kind`: algorithmic flops x the number of
 * bf16 hi/lo products the engine forms (3 or 4 on the tcgen05 engine, 1 on the SIMT engine). */
int msb_profile_read_executed(int kind, double* executed_flops);

#ifdef __cplusplus
}
#endif
#endif /* METASOLVER_B200_H_ */
