// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, M = 128, cta_group::1) by operand source and N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-metasolver_b200/csrc -o gpurun_out/mma_rate_probe scripts/probes/mma_rate_probe.cu
// One CTA per SM, one thread issues NMMA MMAs back to back (K = 16 each), commits, waits; clock64 around.
// Prints clocks per MMA against the nominal floor N/2 (M = 128: 128 x N x 16 MACs at 4096 MAC/clk/SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "msb_ptx.cuh"
using namespace msb;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: SS (A smem tile 128 x 64 sw128, B smem N rows x 64 sw128), 1: TS (A in TMEM)
// AROT: distinct A slabs cycled through (1 or 24); DS: accumulators rotated (1 = one long accumulation chain)
// The issue loop is warp-uniform and unrolled by 24 with compile-time offsets (uniform registers, no election loops).
template <int N, int MODE, int AROT, int DS>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + ((i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tbase, 512); ptx::tmem_relinquish(); }
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
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const uint32_t a_smem = ptx::smem_u32(smem);                  // 4 A tiles of 16 KB
        const uint32_t b_smem = ptx::smem_u32(smem + 64 * 1024);      // B: 3 x 32 KB
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
#pragma unroll
                for (int u = 0; u < 24; ++u) {
                    constexpr int dummy = 0; (void)dummy;
                    const int k = u & 3;
                    const uint32_t d = tbu + (uint32_t)((u % DS) * N);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(b_smem + ((u >> 2) % 3) * 32768 + k * 32, 16, 1024);
                    const uint32_t acc = (it > 0 || u >= DS) ? 1u : 0u;
                    if (MODE == 0) {
                        const uint64_t adesc = ptx::make_smem_desc_sw128(a_smem + ((u >> 2) % (AROT > 4 ? 4 : AROT)) * 16384 + k * 32, 16, 1024);
                        ptx::umma_bf16(d, adesc, bdesc, idesc, acc);
                    } else {
                        umma_ts(d, tbu + 256 + (uint32_t)((u % AROT) * 8), bdesc, idesc, acc);
                    }
                }
            }
            __syncwarp();
        }
        if (leader) ptx::umma_commit(&bar);
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (leader) out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 512); }
}

template <int N, int MODE, int AROT, int DS> double run(long long* dout, int nsm) {
    const int iters = 170;
    const size_t smem = 162 * 1024;
    auto kern = probe<N, MODE, AROT, (DS * N <= 256 ? DS : 1)>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<nsm, 128, smem>>>(iters, dout);
    kern<<<nsm, 128, smem>>>(iters, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    static long long h[256];
    cudaMemcpy(h, dout, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < nsm; ++i) s += (double)h[i];
    return s / nsm / (iters * 24);
}

template <int MODE, int AROT, int DS> void row(const char* name, long long* dout, int grid) {
    printf("%-34s N=64 %6.1f [32]   N=128 %6.1f [64]   N=192 %6.1f [96]   N=256 %6.1f [128]\n", name,
           run<64, MODE, AROT, DS>(dout, grid), run<128, MODE, AROT, DS>(dout, grid), run<192, MODE, AROT, DS>(dout, grid),
           run<256, MODE, AROT, DS>(dout, grid));
}

int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long* dout; cudaMalloc(&dout, 256 * sizeof(long long));
    printf("SMs %d, 4080 MMAs per CTA (M = 128, K = 16, bf16): clocks per MMA  [floor N/2]\n", nsm);
    for (int grid : {1, nsm}) {
        printf("-- grid %d\n", grid);
        row<0, 4, 1>("SS  4 A tiles, 1 accumulator", dout, grid);
        row<0, 4, 2>("SS  4 A tiles, 2 accumulators", dout, grid);
        row<0, 1, 1>("SS  same A tile", dout, grid);
        row<1, 24, 1>("TS 24 A slabs, 1 accumulator", dout, grid);
        row<1, 24, 2>("TS 24 A slabs, 2 accumulators", dout, grid);
        row<1, 1, 1>("TS  same A slab", dout, grid);
    }
    return 0;
}
