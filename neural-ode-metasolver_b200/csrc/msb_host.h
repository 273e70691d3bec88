// Host-side helpers of libmetasolver_b200.so shared by odeblock.cu and blocks.cu (not part of the ABI).
#pragma once
#include "metasolver_b200.h"
#include "msb_internal.h"

namespace msb {

inline size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (1 KB aligned slices).
struct Carver {
    char* base; size_t off, cap;
    Carver(void* p, size_t c) : base((char*)p), off(0), cap(c) {}
    template <typename T> T* take(size_t bytes) {
        size_t o = align_up(off);
        off = o + bytes;
        return (T*)(base + o);
    }
    bool ok() const { return off <= cap; }
};

size_t packed_w_bytes(int engine, int C);
int resolve_engine_shape(int engine, int C, int H, int W);     // MSB_ENGINE_AUTO -> concrete engine, <0 on error
double conv_flops(ConvShape s);
// one convolution launch on `engine` (records profile events when profiling is on)
int run_conv(int engine, const __nv_bfloat16* in, const void* wpacked, const EpiParams& e, ConvShape s, cudaStream_t st);
// (H, W) = the images the packed weights will be used on: the tcgen05 conv form, hence the packing, depends on them
void pack_w(int engine, const float* w, void* out, int C, int transpose, cudaStream_t st, int H, int W);
int wgrad_nparts(int engine, ConvShape s);
// Weight-gradient accumulation over the launches of one backward pass (see odeblock.cu).
struct WgradAcc { float* partial; float* grad_w; int launches; int nparts; };
int run_wgrad(int engine, const __nv_bfloat16* gout, const __nv_bfloat16* in, WgradAcc& acc, ConvShape s, cudaStream_t st);
int wgrad_finish(int engine, WgradAcc& acc, ConvShape s, cudaStream_t st);

// ---- mnist_fused.cu: MNIST ODE block forward as one persistent launch ----
bool mnist_fused_supported(const MsbOdeDesc* d, const MsbMnistParams* mp);
size_t mnist_fused_workspace_bytes();
int launch_mnist_fused_forward(const MsbOdeDesc* d, const float* x, const MsbMnistParams* mp, float* y_out, void* workspace,
                               void* tape, size_t slot_stride, cudaStream_t st);

}  // namespace msb
