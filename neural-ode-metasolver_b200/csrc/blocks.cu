// C ABI of the non-ODE layers of the CIFAR networks: the stem convolution and the strided
// pre-activation residual block.  The residual block's three convolutions (3x3 stride 2, 3x3 stride 1, 1x1
// stride 2) all run on the library's own convolution engines (tcgen05 where the output shape is covered,
// SIMT otherwise) after the space-to-depth re-indexing described in netlayers.cu; nothing here calls cuDNN.
#include "metasolver_b200.h"
#include "msb_host.h"

using namespace msb;

namespace {

struct DownGeom {
    int B, H, W, Ci, Co, Ho, Wo, engine;
    size_t E1, E2;          // elements of the input / output state
    ConvShape shp;          // the stride-1 problems: (B, Ho, Wo, Co)
    size_t wd_bytes;        // 3 space-to-depth weight tensors, fp32 OIHW
};

int down_geom(const MsbDownDesc* d, DownGeom* g) {
    if (!d) { set_error("null descriptor"); return -1; }
    if (d->act != MSB_ACT_GELU_ERF && d->act != MSB_ACT_RELU && d->act != MSB_ACT_NONE) { set_error("unsupported activation %d", d->act); return -1; }
    if (d->batch < 1 || d->height < 2 || d->width < 2 || (d->height & 1) || (d->width & 1)) {
        set_error("strided block: bad input size B=%d H=%d W=%d (H and W must be even)", d->batch, d->height, d->width); return -1;
    }
    if (d->in_channels < 4 || d->in_channels % 4 || d->out_channels != 2 * d->in_channels) {
        set_error("strided block: needs out_channels == 2*in_channels, in_channels %% 4 == 0 (got %d -> %d)", d->in_channels, d->out_channels);
        return -1;
    }
    g->B = d->batch; g->H = d->height; g->W = d->width; g->Ci = d->in_channels; g->Co = d->out_channels;
    g->Ho = g->H / 2; g->Wo = g->W / 2;
    g->engine = resolve_engine_shape(d->engine, g->Co, g->Ho, g->Wo);
    if (g->engine < 0) return -1;
    g->E1 = (size_t)g->B * g->H * g->W * g->Ci;
    g->E2 = (size_t)g->B * g->Ho * g->Wo * g->Co;
    g->shp = ConvShape{g->B, g->Ho, g->Wo, g->Co};
    g->wd_bytes = (size_t)3 * g->Co * g->Co * 9 * sizeof(float);
    return 0;
}

struct DownTape { __nv_bfloat16 *T0, *T1, *Tsc, *Hs; float *G0, *G1; };
size_t down_tape_bytes(const DownGeom& g) { return 4 * align_up(g.E2 * 4) + align_up(g.E1 * 4) + align_up(g.E2 * 4); }
DownTape down_tape(void* base, const DownGeom& g) {
    char* p = (char*)base;
    const size_t q = align_up(g.E2 * 4);
    DownTape t;
    t.T0 = (__nv_bfloat16*)p; t.T1 = (__nv_bfloat16*)(p + q); t.Tsc = (__nv_bfloat16*)(p + 2 * q); t.Hs = (__nv_bfloat16*)(p + 3 * q);
    t.G0 = (float*)(p + 4 * q);
    t.G1 = (float*)(p + 4 * q + align_up(g.E1 * 4));
    return t;
}

}  // namespace

extern "C" {

int msb_stem_forward(const float* x, const float* w, int act, float* y, float* dact_out, int batch, int height, int width,
                     int channels, void* cuda_stream) {
    if (!x || !w || !y || batch < 1 || height < 1 || width < 1) { set_error("msb_stem_forward: bad arguments"); return -1; }
    if (act != MSB_ACT_GELU_ERF && act != MSB_ACT_RELU && act != MSB_ACT_NONE) { set_error("unsupported activation %d", act); return -1; }
    return launch_stem_fwd(x, w, act, y, dact_out, batch, height, width, channels, (cudaStream_t)cuda_stream);
}

size_t msb_stem_backward_workspace_bytes(int channels) {
    return (size_t)stem_wgrad_blocks() * 27 * (size_t)(channels > 0 ? channels : 0) * sizeof(float) + 1024;
}

int msb_stem_backward(const float* grad_y, const float* dact, const float* x, const float* w, float* grad_w, float* grad_x,
                      int batch, int height, int width, int channels, void* workspace, size_t workspace_bytes,
                      void* cuda_stream) {
    if (!grad_y || !dact || !x || !w || batch < 1) { set_error("msb_stem_backward: bad arguments"); return -1; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (grad_w) {
        if (!workspace || workspace_bytes < msb_stem_backward_workspace_bytes(channels)) { set_error("msb_stem_backward: workspace too small"); return -1; }
        if (launch_stem_wgrad(grad_y, dact, x, (float*)workspace, grad_w, batch, height, width, channels, st)) return -1;
    }
    if (grad_x && launch_stem_dgrad(grad_y, dact, w, grad_x, batch, height, width, channels, st)) return -1;
    return 0;
}

size_t msb_downblock_tape_bytes(const MsbDownDesc* d) {
    DownGeom g;
    if (down_geom(d, &g)) return 0;
    return down_tape_bytes(g);
}
size_t msb_downblock_workspace_bytes(const MsbDownDesc* d) {
    DownGeom g;
    if (down_geom(d, &g)) return 0;
    size_t n = align_up(g.wd_bytes) + 4 * align_up(packed_w_bytes(g.engine, g.Co)) + 2 * align_up(g.E2 * 4);
    n += 4 * align_up(g.E2 * 4);           // T0, T1, Tsc, Hs when no tape is recorded
    return n + 4096;
}
size_t msb_downblock_bwd_workspace_bytes(const MsbDownDesc* d) {
    DownGeom g;
    if (down_geom(d, &g)) return 0;
    size_t n = 2 * align_up(g.wd_bytes) + 4 * align_up(packed_w_bytes(g.engine, g.Co));
    n += 2 * align_up(g.E2 * 4);           // Kbar, DP (split)
    n += 3 * align_up(g.E2 * 4);           // gT0, gT1, gTsc (fp32)
    n += align_up((size_t)wgrad_nparts(g.engine, g.shp) * 9 * g.Co * g.Co * 4);
    return n + 4096;
}

int msb_downblock_forward(const MsbDownDesc* d, const float* x, const float* w1, const float* w2, const float* wsc,
                          float* y_out, void* workspace, size_t workspace_bytes, void* tape, size_t tape_bytes,
                          void* cuda_stream) {
    DownGeom g;
    if (down_geom(d, &g)) return -1;
    if (!x || !w1 || !w2 || !wsc || !y_out || !workspace) { set_error("null pointer argument"); return -1; }
    if (workspace_bytes < msb_downblock_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const bool save = d->save_tape != 0;
    if (save && (!tape || tape_bytes < down_tape_bytes(g))) { set_error("tape missing or too small"); return -1; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Carver cv(workspace, workspace_bytes);
    float* Wd = cv.take<float>(g.wd_bytes);
    void* wp[4];
    for (int i = 0; i < 4; ++i) wp[i] = cv.take<char>(packed_w_bytes(g.engine, g.Co));
    float* P = cv.take<float>(g.E2 * 4);
    float* SC = cv.take<float>(g.E2 * 4);
    DownTape t;
    if (save) t = down_tape(tape, g);
    else {
        t.T0 = cv.take<__nv_bfloat16>(g.E2 * 4); t.T1 = cv.take<__nv_bfloat16>(g.E2 * 4);
        t.Tsc = cv.take<__nv_bfloat16>(g.E2 * 4); t.Hs = cv.take<__nv_bfloat16>(g.E2 * 4);
        t.G0 = nullptr; t.G1 = nullptr;
    }
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    const size_t per = (size_t)g.Co * g.Co * 9;
    launch_down_weights_build(w1, wsc, Wd, g.Ci, g.Co, st);
    for (int i = 0; i < 3; ++i) pack_w(g.engine, Wd + i * per, wp[i], g.Co, 0, st, g.shp.H, g.shp.W);
    pack_w(g.engine, w2, wp[3], g.Co, 0, st, g.shp.H, g.shp.W);
    launch_s2d_act_split(x, d->act, t.T0, t.T1, t.Tsc, t.G0, g.B, g.H, g.W, g.Ci, st);
    // shortcut: SC = conv1x1_s2(x)
    EpiParams e = epi_default();
    e.out_f32 = SC;
    if (run_conv(g.engine, t.Tsc, wp[2], e, g.shp, st)) return -1;
    // P = conv3x3_s2(act(x)) as the sum of the two row-phase convolutions; Hs = split(act(P)), G1 = act'(P)
    e = epi_default();
    e.out_f32 = P;
    if (run_conv(g.engine, t.T0, wp[0], e, g.shp, st)) return -1;
    e = epi_default();
    e.base = P; e.act = d->act; e.out_split = t.Hs; e.dact_out = t.G1;
    if (run_conv(g.engine, t.T1, wp[1], e, g.shp, st)) return -1;
    // y = conv2(Hs) + SC
    e = epi_default();
    e.base = SC; e.out_f32 = y_out;
    if (run_conv(g.engine, t.Hs, wp[3], e, g.shp, st)) return -1;
    return check_cuda(cudaGetLastError(), "downblock forward");
}

int msb_downblock_backward(const MsbDownDesc* d, const float* grad_y, const float* w1, const float* w2, const float* wsc,
                           const void* tape, size_t tape_bytes, float* grad_x, float* grad_w1, float* grad_w2,
                           float* grad_wsc, void* workspace, size_t workspace_bytes, void* cuda_stream) {
    DownGeom g;
    if (down_geom(d, &g)) return -1;
    if (!grad_y || !w1 || !w2 || !wsc || !tape || !grad_x || !workspace) { set_error("null pointer argument"); return -1; }
    if (tape_bytes < down_tape_bytes(g)) { set_error("tape too small"); return -1; }
    if (workspace_bytes < msb_downblock_bwd_workspace_bytes(d)) { set_error("workspace too small"); return -1; }
    const int nw = (grad_w1 != nullptr) + (grad_w2 != nullptr) + (grad_wsc != nullptr);
    if (nw != 0 && nw != 3) { set_error("grad_w1, grad_w2, grad_wsc must all be given or all be NULL"); return -1; }
    const bool need_w = nw == 3;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const DownTape t = down_tape(const_cast<void*>(tape), g);
    Carver cv(workspace, workspace_bytes);
    float* Wd = cv.take<float>(g.wd_bytes);
    float* gWd = cv.take<float>(g.wd_bytes);
    void* wt[4];
    for (int i = 0; i < 4; ++i) wt[i] = cv.take<char>(packed_w_bytes(g.engine, g.Co));
    __nv_bfloat16* Kbar = cv.take<__nv_bfloat16>(g.E2 * 4);
    __nv_bfloat16* DP = cv.take<__nv_bfloat16>(g.E2 * 4);
    float* gT[3];
    for (int i = 0; i < 3; ++i) gT[i] = cv.take<float>(g.E2 * 4);
    float* partial = cv.take<float>((size_t)wgrad_nparts(g.engine, g.shp) * 9 * g.Co * g.Co * 4);
    if (!cv.ok()) { set_error("internal: workspace carve overflow"); return -1; }
    const size_t per = (size_t)g.Co * g.Co * 9;
    launch_down_weights_build(w1, wsc, Wd, g.Ci, g.Co, st);
    for (int i = 0; i < 3; ++i) pack_w(g.engine, Wd + i * per, wt[i], g.Co, 1, st, g.shp.H, g.shp.W);
    pack_w(g.engine, w2, wt[3], g.Co, 1, st, g.shp.H, g.shp.W);

    launch_act_split(grad_y, nullptr, ACT_NONE, 1.f, Kbar, nullptr, g.B, g.Ho, g.Wo, g.Co, st);
    auto wgrad_once = [&](const __nv_bfloat16* go, const __nv_bfloat16* in, float* out) {
        WgradAcc acc{partial, out, 0, 0};
        if (run_wgrad(g.engine, go, in, acc, g.shp, st)) return -1;
        return wgrad_finish(g.engine, acc, g.shp, st);
    };
    if (need_w && wgrad_once(Kbar, t.Hs, grad_w2)) return -1;
    // dP = dgrad_W2(gy) * act'(P)
    EpiParams e = epi_default();
    e.mul = t.G1; e.out_split = DP;
    if (run_conv(g.engine, Kbar, wt[3], e, g.shp, st)) return -1;
    if (need_w) {
        if (wgrad_once(DP, t.T0, gWd) || wgrad_once(DP, t.T1, gWd + per) || wgrad_once(Kbar, t.Tsc, gWd + 2 * per)) return -1;
        launch_down_weights_gather(gWd, grad_w1, grad_wsc, g.Ci, g.Co, st);
    }
    const __nv_bfloat16* src[3] = {DP, DP, Kbar};
    for (int i = 0; i < 3; ++i) {
        e = epi_default();
        e.out_f32 = gT[i];
        if (run_conv(g.engine, src[i], wt[i], e, g.shp, st)) return -1;
    }
    launch_d2s_grad(gT[0], gT[1], gT[2], t.G0, grad_x, g.B, g.H, g.W, g.Ci, st);
    return check_cuda(cudaGetLastError(), "downblock backward");
}

}  // extern "C"
