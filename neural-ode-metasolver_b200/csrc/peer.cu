// The one exchange step of the data-parallel path: the gradient all-reduce of a training step
// and the optimizer update that follows it, as ONE kernel over peer memory (NVLink 5 / NVSwitch loads) instead of an
// NCCL all-reduce followed by an update kernel.
//
//   * every rank owns one exchange buffer [header 1 KB | flat fp32 gradient | result] allocated with cudaMalloc, exported with a
//     CUDA IPC handle; the ranks of a node open each other's buffers once (msb_peer_open enables peer access);
//   * a step is one launch per rank: announce "my gradient is complete" to every peer (release store at system scope into
//     the peer's header), wait for all announcements, then every rank reads ALL ranks' gradients element by element in
//     rank order 0..W-1 (the same summation order everywhere: the reduced gradient and hence the updated parameters are
//     bitwise identical on all ranks), scales by 1/W and applies the SGD-momentum / weight-decay update of
//     examples/cifar10/train_and_attack.py:98-99,322 to its own replica (and / or writes the averaged gradient);
//   * the launch ends with the second half of the handshake: the last CTA tells every peer "I have finished reading your
//     gradient" and waits for the same message from all of them, so that the stream-ordered successor of the kernel (the
//     next backward pass) may overwrite the gradient buffer.
// That is the ONE-SHOT form: every rank reads (W-1) x 2.70 MB over NVLink.  Measured (scripts/bench_exchange.py, B200):
// 2 ranks 19.6 us against 34.7 us for ncclAllReduce + update kernel; 8 ranks 49.7 us against 43.9 us (380 GB/s of peer reads
// per rank: NCCL reduces inside the switch and moves each gradient once).  Hence, for three ranks and more, the TWO-SHOT
// form in the same single launch: rank r reduces only slice r of the gradient (reads (W-1)/W x 2.70 MB), stores the
// averaged slice into the result array of EVERY rank's exchange buffer (peer stores), a second handshake says "my slice has
// landed everywhere" (which also tells the peers that this rank no longer reads their gradients), and after it every rank
// updates its replica from its own, now complete, result array.  Same rank-order sum, same bitwise-identical replicas,
// 2 x (W-1)/W x 2.70 MB of NVLink traffic per rank instead of (W-1) x 2.70 MB: 8 ranks 30.0 us (1.56 x NCCL + update).
// The epoch that orders the messages lives in the header (device memory), not in a kernel argument: the launch is
// identical every step and can be captured in a CUDA graph.
// Every wait is bounded (timeout -> header.error, the launch finishes with garbage instead of hanging the GPU).
#include <cuda_runtime.h>

#include "metasolver_b200.h"
#include "msb_internal.h"

namespace msb {
namespace {

struct PeerHeader {                           // first MSB_PEER_HEADER_BYTES of an exchange buffer
    uint32_t ready[MSB_PEER_MAX_RANKS];       // ready[r] = e: rank r's gradient of epoch e is complete       (written by rank r)
    uint32_t done[MSB_PEER_MAX_RANKS];        // done[r]  = e: rank r has finished reading THIS buffer's gradient in e and (two-shot
                                              //               form) its slice of the average has landed in THIS buffer's result array
    uint32_t epoch;                           // last finished epoch of the owner                              (owner only)
    uint32_t ticket;                          // CTAs of the running launch that have finished their reads     (owner only)
    uint32_t error;                           // != 0: the FIRST wait of the owner that timed out (MSB_PEER_ERR_*), sticky
    uint32_t pad;
};
static_assert(sizeof(PeerHeader) <= MSB_PEER_HEADER_BYTES, "header");

struct PeerArgs {
    void* base[MSB_PEER_MAX_RANKS];           // exchange buffers, [rank] = own (peer ones through IPC mappings)
    int world, rank;
    long long offset, n;                      // floats of the gradient region this launch reduces: [offset, offset + n)
    long long result_off;                     // float offset of the result array inside every payload (< 0: none)
    int write_result;                         // one-shot form: also store the average into the own result array
    float* params; float* mom;                // optional: SGD update of the own replica
    float lr, momentum, wd, scale;
    int first;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// gradient loads: system-scope relaxed (never served from a stale non-coherent line of a peer's memory)
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f1(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag has reached epoch e (wrap-safe); false = timed out
__device__ __forceinline__ bool wait_epoch(const uint32_t* flag, uint32_t e, unsigned long long timeout_ns) {
    if ((int32_t)(ld_acquire_sys(flag) - e) >= 0) return true;
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - e) < 0) {
        if (globaltimer_ns() - t0 > timeout_ns) return false;
        __nanosleep(32);
    }
    return true;
}

__device__ __forceinline__ float sgd_update(float g, float* p, float* mom, float lr, float momentum, float wd, int first) {
    // same operations, same order as sgd_step_kernel (train_aux.cu)
    const float w = *p;
    if (wd != 0.f) g = __fadd_rn(g, __fmul_rn(wd, w));
    float b = g;
    if (momentum != 0.f) {
        b = first ? g : __fadd_rn(__fmul_rn(momentum, *mom), g);
        *mom = b;
    }
    *p = __fsub_rn(w, __fmul_rn(lr, b));
    return g;
}

template <int W>
__global__ void __launch_bounds__(256) peer_allreduce_sgd_kernel(PeerArgs a) {
    PeerHeader* own = reinterpret_cast<PeerHeader*>(a.base[a.rank]);
    const uint32_t e = own->epoch + 1;        // written by the last CTA of the previous launch (stream order)
    const int tid = threadIdx.x;
    __shared__ int s_last;

    // ---- handshake 1: gradients complete everywhere ----
    if (blockIdx.x == 0 && tid < a.world) {
        __threadfence_system();               // the gradient was written by earlier kernels of this stream
        st_release_sys(&reinterpret_cast<PeerHeader*>(a.base[tid])->ready[a.rank], e);
    }
    if (tid < a.world && !wait_epoch(&own->ready[tid], e, a.timeout_ns)) atomicCAS(&own->error, 0u, (uint32_t)MSB_PEER_ERR_READY_TIMEOUT);
    __syncthreads();

    // ---- reduce in rank order + update ----
    const bool do_sgd = a.params != nullptr;
    float* const avg_out = a.write_result ? reinterpret_cast<float*>(reinterpret_cast<char*>(a.base[a.rank]) + MSB_PEER_HEADER_BYTES) + a.result_off + a.offset
                                          : nullptr;
    const bool vec = (a.offset & 3) == 0 && (a.result_off < 0 || (a.result_off & 3) == 0) &&
                     (reinterpret_cast<uintptr_t>(a.params) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.mom) & 15) == 0;
    const long long n4 = vec ? a.n / 4 : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n4; i += stride) {
        float4 g[W];
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r < a.world)
                g[r] = ld_sys_f4(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(a.base[r]) + MSB_PEER_HEADER_BYTES) + a.offset / 4 + i);
        float4 s = g[0];
#pragma unroll
        for (int r = 1; r < W; ++r)
            if (r < a.world) {
                s.x = __fadd_rn(s.x, g[r].x); s.y = __fadd_rn(s.y, g[r].y);
                s.z = __fadd_rn(s.z, g[r].z); s.w = __fadd_rn(s.w, g[r].w);
            }
        if (a.scale != 1.f) { s.x = __fmul_rn(s.x, a.scale); s.y = __fmul_rn(s.y, a.scale); s.z = __fmul_rn(s.z, a.scale); s.w = __fmul_rn(s.w, a.scale); }
        if (avg_out) reinterpret_cast<float4*>(avg_out)[i] = s;
        if (do_sgd) {
            float4 p = reinterpret_cast<float4*>(a.params)[i];
            float4 m = (a.momentum != 0.f && !a.first) ? reinterpret_cast<float4*>(a.mom)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            sgd_update(s.x, &p.x, &m.x, a.lr, a.momentum, a.wd, a.first);
            sgd_update(s.y, &p.y, &m.y, a.lr, a.momentum, a.wd, a.first);
            sgd_update(s.z, &p.z, &m.z, a.lr, a.momentum, a.wd, a.first);
            sgd_update(s.w, &p.w, &m.w, a.lr, a.momentum, a.wd, a.first);
            reinterpret_cast<float4*>(a.params)[i] = p;
            if (a.momentum != 0.f) reinterpret_cast<float4*>(a.mom)[i] = m;
        }
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + tid; i < a.n; i += stride) {      // tail / unaligned runs
        float s = 0.f;
        for (int r = 0; r < a.world; ++r) {
            const float g = ld_sys_f1(reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.base[r]) + MSB_PEER_HEADER_BYTES) + a.offset + i);
            s = r == 0 ? g : __fadd_rn(s, g);
        }
        if (a.scale != 1.f) s = __fmul_rn(s, a.scale);
        if (avg_out) avg_out[i] = s;
        if (do_sgd) sgd_update(s, a.params + i, a.mom ? a.mom + i : nullptr, a.lr, a.momentum, a.wd, a.first);
    }

    // ---- handshake 2: nobody reads a gradient buffer any more when the launches end ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&own->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (tid < a.world) {
        st_release_sys(&reinterpret_cast<PeerHeader*>(a.base[tid])->done[a.rank], e);
        if (!wait_epoch(&own->done[tid], e, a.timeout_ns)) atomicCAS(&own->error, 0u, (uint32_t)MSB_PEER_ERR_DONE_TIMEOUT);
    }
    __syncthreads();
    if (tid == 0) {
        own->ticket = 0;
        own->epoch = e;
    }
}

// TWO-SHOT form (see the file comment).  Host guarantees: offset, result_off multiples of 4, params / mom 16-byte aligned,
// gridDim.x CTAs co-resident (the second handshake is a grid-wide barrier).
template <int W>
__global__ void __launch_bounds__(256) peer_twoshot_sgd_kernel(PeerArgs a) {
    PeerHeader* own = reinterpret_cast<PeerHeader*>(a.base[a.rank]);
    const uint32_t e = own->epoch + 1;
    const int tid = threadIdx.x;
    __shared__ int s_last;

    // ---- handshake 1: gradients complete everywhere ----
    if (blockIdx.x == 0 && tid < a.world) {
        __threadfence_system();
        st_release_sys(&reinterpret_cast<PeerHeader*>(a.base[tid])->ready[a.rank], e);
    }
    if (tid < a.world && !wait_epoch(&own->ready[tid], e, a.timeout_ns)) atomicCAS(&own->error, 0u, (uint32_t)MSB_PEER_ERR_READY_TIMEOUT);
    __syncthreads();

    // ---- reduce-scatter: this rank's slice, rank-order sum, the average stored into every rank's result array ----
    const long long n4 = a.n / 4;
    const long long per = (n4 + a.world - 1) / a.world;
    const long long lo = (long long)a.rank * per, hi = lo + per < n4 ? lo + per : n4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long g4 = a.offset / 4, r4 = (a.result_off + a.offset) / 4;
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + tid; i < hi; i += stride) {
        float4 g[W];
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r < a.world)
                g[r] = ld_sys_f4(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(a.base[r]) + MSB_PEER_HEADER_BYTES) + g4 + i);
        float4 s = g[0];
#pragma unroll
        for (int r = 1; r < W; ++r)
            if (r < a.world) {
                s.x = __fadd_rn(s.x, g[r].x); s.y = __fadd_rn(s.y, g[r].y);
                s.z = __fadd_rn(s.z, g[r].z); s.w = __fadd_rn(s.w, g[r].w);
            }
        if (a.scale != 1.f) { s.x = __fmul_rn(s.x, a.scale); s.y = __fmul_rn(s.y, a.scale); s.z = __fmul_rn(s.z, a.scale); s.w = __fmul_rn(s.w, a.scale); }
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r < a.world)
                reinterpret_cast<float4*>(reinterpret_cast<char*>(a.base[r]) + MSB_PEER_HEADER_BYTES)[r4 + i] = s;
    }
    if (a.rank == 0 && blockIdx.x == 0 && tid < (int)(a.n - n4 * 4)) {          // the <= 3 floats after the last float4: rank 0
        const long long i = n4 * 4 + tid;
        float s = 0.f;
        for (int r = 0; r < a.world; ++r) {
            const float g = ld_sys_f1(reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.base[r]) + MSB_PEER_HEADER_BYTES) + a.offset + i);
            s = r == 0 ? g : __fadd_rn(s, g);
        }
        if (a.scale != 1.f) s = __fmul_rn(s, a.scale);
        for (int r = 0; r < a.world; ++r)
            (reinterpret_cast<float*>(reinterpret_cast<char*>(a.base[r]) + MSB_PEER_HEADER_BYTES) + a.result_off + a.offset)[i] = s;
    }

    // ---- handshake 2: "my slice has landed everywhere" (implies: I no longer read anybody's gradient) ----
    __threadfence_system();                   // this thread's peer stores are performed system-wide ...
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();               // ... and so are the CTA's, before the ticket (cumulativity through the barrier)
        s_last = atomicAdd(&own->ticket, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (s_last) {                             // every CTA of this rank has stored its part
        if (tid < a.world) st_release_sys(&reinterpret_cast<PeerHeader*>(a.base[tid])->done[a.rank], e);
        if (tid == 0) {
            own->ticket = 0;
            own->epoch = e;                   // nobody of this launch reads it again
        }
    }
    if (tid < a.world && !wait_epoch(&own->done[tid], e, a.timeout_ns)) atomicCAS(&own->error, 0u, (uint32_t)MSB_PEER_ERR_DONE_TIMEOUT);
    __syncthreads();
    if (!a.params) return;                    // the own result array now holds the whole averaged gradient

    // ---- update the replica from the own result array (written by the peers: system-scope loads) ----
    const float* res = reinterpret_cast<const float*>(reinterpret_cast<const char*>(a.base[a.rank]) + MSB_PEER_HEADER_BYTES) + a.result_off + a.offset;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n4; i += stride) {
        const float4 s = ld_sys_f4(reinterpret_cast<const float4*>(res) + i);
        float4 p = reinterpret_cast<float4*>(a.params)[i];
        float4 m = (a.momentum != 0.f && !a.first) ? reinterpret_cast<float4*>(a.mom)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        sgd_update(s.x, &p.x, &m.x, a.lr, a.momentum, a.wd, a.first);
        sgd_update(s.y, &p.y, &m.y, a.lr, a.momentum, a.wd, a.first);
        sgd_update(s.z, &p.z, &m.z, a.lr, a.momentum, a.wd, a.first);
        sgd_update(s.w, &p.w, &m.w, a.lr, a.momentum, a.wd, a.first);
        reinterpret_cast<float4*>(a.params)[i] = p;
        if (a.momentum != 0.f) reinterpret_cast<float4*>(a.mom)[i] = m;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + tid; i < a.n; i += stride)
        sgd_update(ld_sys_f1(res + i), a.params + i, a.mom ? a.mom + i : nullptr, a.lr, a.momentum, a.wd, a.first);
}

template <typename K>
int coresident_ctas(K kern) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * num_sms();
}

}  // namespace
}  // namespace msb

using namespace msb;

extern "C" {

int msb_peer_alloc(size_t payload_bytes, void** base, unsigned char* handle) {
    if (!base || !handle) { set_error("msb_peer_alloc: null argument"); return -1; }
    void* p = nullptr;
    const size_t bytes = MSB_PEER_HEADER_BYTES + (payload_bytes + 15) / 16 * 16;
    if (check_cuda(cudaMalloc(&p, bytes), "msb_peer_alloc: cudaMalloc")) return -1;
    if (check_cuda(cudaMemset(p, 0, bytes), "msb_peer_alloc: cudaMemset")) { cudaFree(p); return -1; }
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == MSB_PEER_HANDLE_BYTES, "IPC handle size");
    if (check_cuda(cudaIpcGetMemHandle(&h, p), "msb_peer_alloc: cudaIpcGetMemHandle")) { cudaFree(p); return -1; }
    if (check_cuda(cudaDeviceSynchronize(), "msb_peer_alloc: sync")) { cudaFree(p); return -1; }
    memcpy(handle, &h, sizeof(h));
    *base = p;
    return 0;
}

int msb_peer_open(const unsigned char* handle, void** base) {
    if (!base || !handle) { set_error("msb_peer_open: null argument"); return -1; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    if (check_cuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "msb_peer_open: cudaIpcOpenMemHandle")) return -1;
    *base = p;
    return 0;
}

int msb_peer_close(void* base) { return base ? check_cuda(cudaIpcCloseMemHandle(base), "msb_peer_close") : 0; }
int msb_peer_free(void* base) { return base ? check_cuda(cudaFree(base), "msb_peer_free") : 0; }

int msb_peer_status(const void* own_base, unsigned* error_word, unsigned* epoch) {
    if (!own_base) { set_error("msb_peer_status: null buffer"); return -1; }
    PeerHeader h;
    if (check_cuda(cudaMemcpy(&h, own_base, sizeof(h), cudaMemcpyDeviceToHost), "msb_peer_status: copy")) return -1;
    if (error_word) *error_word = h.error;
    if (epoch) *epoch = h.epoch;
    return 0;
}

int msb_peer_allreduce_sgd(void* const* bases, int world, int rank, int64_t offset, int64_t n, int64_t result_offset,
                           int write_result, float* params, float* momentum_buf, float lr, float momentum, float weight_decay,
                           float grad_scale, int first_step, unsigned timeout_ms, void* cuda_stream) {
    if (!bases || world < 1 || world > MSB_PEER_MAX_RANKS || rank < 0 || rank >= world || offset < 0 || n < 0) {
        set_error("msb_peer_allreduce_sgd: bad arguments (world=%d rank=%d offset=%lld n=%lld; at most %d ranks)", world, rank,
                  (long long)offset, (long long)n, MSB_PEER_MAX_RANKS);
        return -1;
    }
    if (params && momentum != 0.f && !momentum_buf) { set_error("msb_peer_allreduce_sgd: momentum without a momentum buffer"); return -1; }
    if (write_result && result_offset < 0) { set_error("msb_peer_allreduce_sgd: write_result without a result array"); return -1; }
    if (!params && !write_result) { set_error("msb_peer_allreduce_sgd: neither parameters to update nor a result to write"); return -1; }
    PeerArgs a;
    for (int r = 0; r < MSB_PEER_MAX_RANKS; ++r) {
        a.base[r] = r < world ? bases[r] : nullptr;
        if (r < world && !bases[r]) { set_error("msb_peer_allreduce_sgd: exchange buffer of rank %d is null", r); return -1; }
    }
    a.world = world; a.rank = rank; a.offset = offset; a.n = n;
    a.result_off = result_offset; a.write_result = write_result ? 1 : 0;
    a.params = params; a.mom = params ? momentum_buf : nullptr;
    a.lr = lr; a.momentum = momentum; a.wd = weight_decay; a.scale = grad_scale; a.first = first_step;
    a.timeout_ns = (unsigned long long)(timeout_ms ? timeout_ms : 10000u) * 1000000ull;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long work = std::max<long long>((n + 3) / 4, 1);
    const int form = tune_get(TUNE_PEER_FORM);          // 0 = two-shot from three ranks on, 1 = always one-shot, 2 = always two-shot
    const bool aligned = result_offset >= 0 && (offset & 3) == 0 && (result_offset & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(params) & 15) == 0 && (reinterpret_cast<uintptr_t>(momentum_buf) & 15) == 0;
    const bool two_shot = aligned && n >= 4 && (form == 2 || (form == 0 && world >= 3));
    if (two_shot) {
        int cap;
        if (world <= 2) cap = coresident_ctas(peer_twoshot_sgd_kernel<2>);
        else if (world <= 4) cap = coresident_ctas(peer_twoshot_sgd_kernel<4>);
        else if (world <= 8) cap = coresident_ctas(peer_twoshot_sgd_kernel<8>);
        else cap = coresident_ctas(peer_twoshot_sgd_kernel<16>);
        const int grid = (int)std::min<long long>((work + 255) / 256, (long long)cap);
        if (world <= 2) peer_twoshot_sgd_kernel<2><<<grid, 256, 0, st>>>(a);
        else if (world <= 4) peer_twoshot_sgd_kernel<4><<<grid, 256, 0, st>>>(a);
        else if (world <= 8) peer_twoshot_sgd_kernel<8><<<grid, 256, 0, st>>>(a);
        else peer_twoshot_sgd_kernel<16><<<grid, 256, 0, st>>>(a);
    } else {
        const int grid = (int)std::min<long long>((work + 255) / 256, 2LL * num_sms());
        if (world <= 2) peer_allreduce_sgd_kernel<2><<<grid, 256, 0, st>>>(a);
        else if (world <= 4) peer_allreduce_sgd_kernel<4><<<grid, 256, 0, st>>>(a);
        else if (world <= 8) peer_allreduce_sgd_kernel<8><<<grid, 256, 0, st>>>(a);
        else peer_allreduce_sgd_kernel<16><<<grid, 256, 0, st>>>(a);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "peer_allreduce_sgd launch");
}

}  // extern "C"
