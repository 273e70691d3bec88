// Microbenchmark: issue rate of tcgen05.mma.cta_group::2 (kind::f16, bf16, M = 256 over a CTA pair, both operands in
// shared memory, each CTA holding half of the N rows of B) by N, and of the N = 256 / N = 128 mix the C = 128 convolution
// issues per k-step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-metasolver_b200/csrc -o build/mma_2sm_probe scripts/probes/mma_2sm_probe.cu
// One pair per two SMs, the leader's elected lane issues MMAs back to back (K = 16 each), commits, waits; clock64 around.
// Prints clocks per MMA against the nominal floor N/2 (each SM forms its 128 x N x 16 part at 4096 MAC/clk).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "msb_ptx.cuh"
using namespace msb;

// MIX 0: every MMA has N columns; MIX 1: alternate N = 256 (A tile 0) and N = 128 (A tile 1), as conv3x3_tcp2 does
// COMMIT_EVERY: MMAs between tcgen05.commit's to a scratch barrier nobody waits on (0 = only the final commit)
template <int N, int MIX, int COMMIT_EVERY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, scratch_bar;
    __shared__ uint32_t tbase;
    const uint32_t rank = ptx::cluster_ctarank();
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + ((i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&scratch_bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc2(&tbase, 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::cluster_sync(); ptx::tc_fence_after();
    const uint32_t tb = tbase;
    long long t0 = 0;
    if (rank == 0 && threadIdx.x < 32) {
        constexpr uint32_t idesc_n = ptx::make_idesc_bf16(256, N, 0, 0);
        constexpr uint32_t idesc_256 = ptx::make_idesc_bf16(256, 256, 0, 0);
        constexpr uint32_t idesc_128 = ptx::make_idesc_bf16(256, 128, 0, 0);
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const uint32_t a_smem = ptx::smem_u32(smem);                  // 4 A tiles of 16 KB (128 rows x 128 B)
        const uint32_t b_smem = ptx::smem_u32(smem + 64 * 1024);      // B halves: 3 x 32 KB
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
#pragma unroll
                for (int u = 0; u < 24; ++u) {
                    const int k = u & 3;
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(b_smem + ((u >> 2) % 3) * 32768 + k * 32, 16, 1024);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(a_smem + ((u >> 2) & 3) * 16384 + k * 32, 16, 1024);
                    const uint32_t acc = (it > 0 || u > 0) ? 1u : 0u;
                    if (MIX == 0) {
                        ptx::umma_bf16_2sm(tbu, adesc, bdesc, idesc_n, acc);
                    } else {
                        const uint64_t adesc2 = ptx::make_smem_desc_sw128(a_smem + (((u >> 2) + 1) & 3) * 16384 + k * 32, 16, 1024);
                        const uint64_t bdesc2 = ptx::make_smem_desc_sw128(b_smem + ((u >> 2) % 3) * 32768 + 16384 + k * 32, 16, 1024);
                        ptx::umma_bf16_2sm(tbu, adesc, bdesc, idesc_256, acc);
                        ptx::umma_bf16_2sm(tbu, adesc2, bdesc2, idesc_128, 1u);
                    }
                    if (COMMIT_EVERY > 0 && (u + 1) % COMMIT_EVERY == 0) ptx::umma_commit_2sm(&scratch_bar);
                }
            }
            __syncwarp();
        }
        if (leader) ptx::umma_commit_2sm(&bar);
        __syncwarp();
    }
    if (threadIdx.x < 32) {
        ptx::mbar_wait(&bar, 0);
        ptx::tc_fence_after();
        if (rank == 0 && threadIdx.x == 0) out[blockIdx.x >> 1] = clock64() - t0;
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::cluster_sync();
    if (threadIdx.x < 32) ptx::tmem_dealloc2(tb, 512);
}

template <int N, int MIX, int CE>
static double run(int grid, int iters) {
    long long* d;
    cudaMalloc(&d, sizeof(long long) * 256);
    const size_t smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(probe2<N, MIX, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe2<N, MIX, CE><<<grid, 128, smem>>>(4, d);           // warm-up
    cudaDeviceSynchronize();
    probe2<N, MIX, CE><<<grid, 128, smem>>>(iters, d);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return -1; }
    long long h[256];
    cudaMemcpy(h, d, sizeof(long long) * (grid / 2), cudaMemcpyDeviceToHost);
    cudaFree(d);
    long long mx = 0;
    for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
    return (double)mx / ((double)iters * 24 * (MIX ? 2 : 1));     // SM clocks per MMA (slowest pair)
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000;
    printf("SMs %d, cta_group::2 MMAs (M = 256, K = 16, bf16, SS), %d per pair: SM clocks per MMA (clock64 in the issuing warp, slowest pair)  [floor N/2]\n", sms, iters * 24);
    for (int grid : {2, (sms / 2) * 2}) {
        printf("-- %d pairs\n", grid / 2);
        printf("one shape, final commit only   N=64 %6.1f [32]   N=128 %6.1f [64]   N=192 %6.1f [96]   N=256 %6.1f [128]\n",
               run<64, 0, 0>(grid, iters), run<128, 0, 0>(grid, iters), run<192, 0, 0>(grid, iters), run<256, 0, 0>(grid, iters));
        printf("one shape, commit every 8      N=64 %6.1f [32]   N=128 %6.1f [64]   N=192 %6.1f [96]   N=256 %6.1f [128]\n",
               run<64, 0, 8>(grid, iters), run<128, 0, 8>(grid, iters), run<192, 0, 8>(grid, iters), run<256, 0, 8>(grid, iters));
        printf("mix N=256 + N=128 (conv3x3_tcp2's k-step): %6.1f per MMA [96], commit every 4 k-steps: %6.1f\n",
               run<256, 1, 0>(grid, iters), run<256, 1, 4>(grid, iters));
    }
    return 0;
}
