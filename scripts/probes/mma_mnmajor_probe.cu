// Probe: rate of the weight gradient's MMA shape -- M = 128, N = 192, K = 16, bf16, BOTH operands MN-major in shared memory --
// by the placement of the three 64-channel N atoms: LBO = 8192 B (three image rows, round-1 geometry), LBO = 128 B (three
// pixels of one staged line, overlapping; option wgrad_htaps), and by the alignment of the operand start.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-metasolver_b200/csrc -o build/mma_mnmajor_probe scripts/probes/mma_mnmajor_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "msb_ptx.cuh"
using namespace msb;

// A_LBO: stride between the two 64-channel M atoms; B_LBO: between the three N atoms; B_OFF: byte offset of the B start
template <int A_LBO, int B_LBO, int B_OFF, int B_PITCH>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + ((i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tbase, 512); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tb = tbase;
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 192, 1, 1);
        const uint32_t tbu = __shfl_sync(0xffffffffu, tb, 0);
        const uint32_t a_smem = ptx::smem_u32(smem);                  // 32 KB: [4 rows][2 planes][32 px][128 B]
        const uint32_t b_smem = ptx::smem_u32(smem + 32 * 1024);      // up to 64 KB
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {                        // one tile of the C = 64 weight gradient: 4 rows x 2 wg x 2 planes
                    const int rho = u >> 2, wg = (u >> 1) & 1, pb = u & 1;
                    const uint64_t adesc = ptx::make_smem_desc_sw128(a_smem + rho * 8192 + wg * 2048, A_LBO, 1024);
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(b_smem + B_OFF + rho * 2 * B_PITCH + pb * B_PITCH + wg * 2048, B_LBO, 1024);
                    ptx::umma_bf16(tbu, adesc, bdesc, idesc, (it > 0 || u > 0) ? 1u : 0u);
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

template <int A_LBO, int B_LBO, int B_OFF, int B_PITCH> void run(const char* name, long long* dout, int nsm) {
    const size_t smem = 162 * 1024;
    auto kern = probe<A_LBO, B_LBO, B_OFF, B_PITCH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 400;
    kern<<<nsm, 128, smem>>>(iters, dout);
    kern<<<nsm, 128, smem>>>(iters, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    static long long h[256];
    cudaMemcpy(h, dout, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < nsm; ++i) s += (double)h[i];
    printf("%-86s %6.1f clk per MMA  [96]\n", name, s / nsm / (iters * 16.0));
}

int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long* dout; cudaMalloc(&dout, 256 * sizeof(long long));
    printf("M = 128 (two 64-channel atoms, LBO 4096), N = 192 (three atoms), K = 16, both operands MN-major, SS; all %d SMs\n", nsm);
    run<4096, 8192, 0, 4096>("B atoms = three image rows (LBO 8192), lines of 32 px, aligned start", dout, nsm);
    run<4096, 8192, 128, 4096>("B atoms = three image rows (LBO 8192), start + 128 B", dout, nsm);
    run<4096, 128, 0, 4352>("B atoms = three pixels of one line (LBO 128, overlapping), lines of 34 px", dout, nsm);
    run<4096, 128, 0, 4096>("B atoms = three pixels of one line (LBO 128, overlapping), lines of 32 px", dout, nsm);
    run<4096, 1024, 0, 4352>("B atoms one 8-pixel group apart (LBO 1024, overlapping), lines of 34 px", dout, nsm);
    run<4096, 2048, 0, 4352>("B atoms 16 pixels apart (LBO 2048, disjoint within a line)", dout, nsm);
    run<16384, 128, 0, 4352>("as (LBO 128, 34 px) with the two M atoms 16 KB apart (C = 128-style A)", dout, nsm);
    return 0;
}
