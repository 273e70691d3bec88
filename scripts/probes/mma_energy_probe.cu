// Probe: sustained tcgen05.mma rate under the board's power cap by MMA shape -- is an M = 64 MMA cheaper (in energy) than
// an M = 128 one?  All SMs issue TS MMAs (A in TMEM, B in shared memory, N = 64, K = 16, bf16) back to back for a few
// seconds per shape; printed: MMAs per second per SM (event-timed), clocks per MMA (clock64), the implied SM clock.
// A shape that sustains more MMAs per second under the same cap costs less energy per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-metasolver_b200/csrc -o build/mma_energy_probe scripts/probes/mma_energy_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "msb_ptx.cuh"
using namespace msb;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MIX: 0 = every MMA has M rows; 1 = alternate M = 128 and M = 64 (the three-product k-step: [W_hi;W_lo] x X_hi, W_hi x X_lo)
template <int M, int N, int MIX>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + ((i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tbase, 512); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tb = tbase;
    {
        uint32_t r[16];
        for (int j = 0; j < 16; ++j) r[j] = 0x3c003c00u + threadIdx.x + j;
        const uint32_t lane_addr = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
        for (int c = 256; c < 512; c += 16)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                         ::"r"(tb + lane_addr + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc_m = ptx::make_idesc_bf16(M, N, 0, 0);
        constexpr uint32_t idesc_64 = ptx::make_idesc_bf16(64, N, 0, 0);
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const uint32_t b_smem = ptx::smem_u32(smem);
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
#pragma unroll
                for (int u = 0; u < 24; ++u) {
                    const int k = u & 3;
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(b_smem + ((u >> 2) % 3) * 32768 + k * 32, 16, 1024);
                    const uint32_t acc = (it > 0 || u > 0) ? 1u : 0u;
                    const uint32_t a = tbu + 256 + (uint32_t)(u * 8);
                    umma_ts(tbu, a, bdesc, (MIX && (u & 1)) ? idesc_64 : idesc_m, acc);
                }
            }
            __syncwarp();
        }
        if (leader) ptx::umma_commit(&bar);
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        if (leader) out[blockIdx.x] = clock64() - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 512); }
}

template <int M, int N, int MIX> void run(const char* name, long long* dout, int nsm, double seconds) {
    const size_t smem = 98 * 1024;
    auto kern = probe<M, N, MIX>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 20000;                                   // 480k MMAs per launch: ~10 ms
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 20; ++w) kern<<<nsm, 128, smem>>>(iters, dout);          // warm up into the power cap
    cudaDeviceSynchronize();
    int launches = 0; float ms = 0;
    cudaEventRecord(e0);
    do {
        for (int j = 0; j < 20; ++j) kern<<<nsm, 128, smem>>>(iters, dout);
        launches += 20;
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < seconds * 1e3);
    static long long h[256];
    cudaMemcpy(h, dout, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double clk = 0; for (int i = 0; i < nsm; ++i) clk += (double)h[i];
    clk /= nsm;
    const double mmas = (double)launches * iters * 24;
    printf("%-46s %8.2f M MMAs/s per SM   %6.1f clk per MMA   SM clock ~%.0f MHz   (%d launches, %.1f s)\n", name,
           mmas / (ms * 1e-3) / 1e6, clk / (iters * 24.0), mmas / (ms * 1e-3) * (clk / (iters * 24.0)) / 1e6, launches, ms * 1e-3);
}

int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long* dout; cudaMalloc(&dout, 256 * sizeof(long long));
    printf("all %d SMs, TS MMAs (A in TMEM), K = 16, bf16, sustained for ~3 s per shape\n", nsm);
    for (int rep = 0; rep < 2; ++rep) {
        run<128, 64, 0>("M = 128, N = 64", dout, nsm, 3.0);
        run<64, 64, 0>("M =  64, N = 64", dout, nsm, 3.0);
        run<128, 64, 1>("M = 128 / M = 64 alternating, N = 64", dout, nsm, 3.0);
        run<128, 128, 0>("M = 128, N = 128", dout, nsm, 3.0);
    }
    return 0;
}
