// Microbenchmark 2: what the per-row synchronisation around a batch of tcgen05.mma costs on the issuing warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-metasolver_b200/csrc -o build/mma_sync_probe scripts/probes/mma_sync_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "msb_ptx.cuh"
using namespace msb;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect() {
    uint32_t p;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(p));
    return p != 0;
}

// out[0] issue clocks of 36 MMAs, out[1] until complete, out[2] issue of 144, out[3] complete of 144,
// out[4] commit + wait with an empty pipe, out[5] already-complete wait (32 lanes), out[6] already-complete wait (1 lane + syncwarp)
// out[7] steady-state clocks per row: [36 MMAs, commit] with a wait on the commit of two rows earlier, all lanes waiting
// out[8] the same, one lane waiting      out[9] the same without any wait (issue + commit only)
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + ((i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tbase, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tb = tbase;
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const uint32_t b_smem = ptx::smem_u32(smem);
        const bool leader = elect();
        auto issue36 = [&](uint32_t d) {
            if (leader) {
#pragma unroll
                for (int u = 0; u < 36; ++u) {
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(b_smem + (u / 12) * 8192 + ((u >> 2) % 3) * 8192 + (u & 3) * 32, 16, 1024);
                    umma_ts(d, tbu + 192 + (uint32_t)(u * 8), bdesc, idesc, u ? 1u : 0u);
                }
            }
            __syncwarp();
        };
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // warm-up
        issue36(tbu); if (leader) ptx::umma_commit(&bar[0]); __syncwarp(); ptx::mbar_wait(&bar[0], ph[0]); ph[0] ^= 1;
        long long t0 = clock64();
        issue36(tbu);
        long long t1 = clock64();
        if (leader) ptx::umma_commit(&bar[0]); __syncwarp(); ptx::mbar_wait(&bar[0], ph[0]); ph[0] ^= 1;
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
        t0 = clock64();
        issue36(tbu); issue36(tbu + 64); issue36(tbu + 128); issue36(tbu);
        t1 = clock64();
        if (leader) ptx::umma_commit(&bar[0]); __syncwarp(); ptx::mbar_wait(&bar[0], ph[0]); ph[0] ^= 1;
        t2 = clock64();
        out[2] = t1 - t0; out[3] = t2 - t0;
        t0 = clock64();
        for (int r = 0; r < 16; ++r) { if (leader) ptx::umma_commit(&bar[0]); __syncwarp(); ptx::mbar_wait(&bar[0], ph[0]); ph[0] ^= 1; }
        out[4] = (clock64() - t0) / 16;
        // already-complete waits: bar[1] completed once (phase 0 done)
        if (leader) ptx::umma_commit(&bar[1]); __syncwarp(); ptx::mbar_wait(&bar[1], 0);
        t0 = clock64();
        for (int r = 0; r < 64; ++r) { ptx::mbar_wait(&bar[1], 0); ptx::tc_fence_after(); }
        out[5] = (clock64() - t0) / 64;
        t0 = clock64();
        for (int r = 0; r < 64; ++r) { if (leader) ptx::mbar_wait(&bar[1], 0); __syncwarp(); ptx::tc_fence_after(); }
        out[6] = (clock64() - t0) / 64;
        // steady state rows: 3 accumulators, commit to bar[2 + row % 3], wait for the commit of row - 2 before row
        for (int mode = 0; mode < 3; ++mode) {
            uint32_t p3[3] = {ph[2], ph[3], ph[4]};
            const int ROWS = 96;
            t0 = clock64();
            for (int r = 0; r < ROWS; ++r) {
                if (mode < 2 && r >= 2) {
                    const int b = (r - 2) % 3;
                    if (mode == 0) ptx::mbar_wait(&bar[2 + b], p3[b]);
                    else { if (leader) ptx::mbar_wait(&bar[2 + b], p3[b]); __syncwarp(); }
                    ptx::tc_fence_after();
                    p3[b] ^= 1;
                }
                issue36(tbu + (uint32_t)((r % 3) * 64));
                if (leader) ptx::umma_commit(&bar[2 + r % 3]);
                __syncwarp();
            }
            // drain
            if (mode < 2) {
                for (int r = ROWS - 2; r < ROWS; ++r) { const int b = r % 3; ptx::mbar_wait(&bar[2 + b], p3[b]); p3[b] ^= 1; }
            } else {
                for (int r = 0; r < ROWS; ++r) { const int b = r % 3; p3[b] ^= 1; }
                if (leader) ptx::umma_commit(&bar[5]); __syncwarp(); ptx::mbar_wait(&bar[5], ph[5]); ph[5] ^= 1;
            }
            out[7 + mode] = (clock64() - t0) / ROWS;
            ph[2] = p3[0]; ph[3] = p3[1]; ph[4] = p3[2];
        }
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 512); }
}

int main() {
    long long* dout; cudaMalloc(&dout, 16 * sizeof(long long));
    const size_t smem = 66 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, smem>>>(dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[16]; cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("36 TS MMAs (N = 64, floor 1152 clk): issued after %lld clk, complete (commit + wait) after %lld clk\n", h[0], h[1]);
    printf("144 MMAs (floor 4608): issued after %lld, complete after %lld\n", h[2], h[3]);
    printf("commit + wait, empty pipe: %lld clk\n", h[4]);
    printf("already-complete wait: all 32 lanes %lld clk, one lane + syncwarp %lld clk\n", h[5], h[6]);
    printf("steady state per row [wait(row-2), 36 MMAs, commit]: all lanes wait %lld clk, one lane waits %lld clk, no waits %lld clk\n", h[7], h[8], h[9]);
    return 0;
}
