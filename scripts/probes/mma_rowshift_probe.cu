// Probe: does tcgen05.mma honour a shared-memory operand whose START is offset by d x 128 bytes (d rows) inside a
// 128-byte-swizzle atom (8 rows x 128 B)?  If it does, a horizontally shifted 3x3-convolution tap can read the SAME
// staged tile with another start address instead of its own shifted copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-metasolver_b200/csrc -o build/mma_rowshift_probe scripts/probes/mma_rowshift_probe.cu
// Test 1 (K-major A, rows = M = pixels; conv3x3_tcp / tcp2): A tile of 144 rows written with the swizzle of its
//   1024-aligned base, A[row][k] = f(row, k); B = one-hot (B[n][k] = [k == n]); D[m][n] must equal A[m + d][n].
// Test 2 (MN-major B, rows = K = pixels, 64 channels per row; wgrad3x3_tc): B[kpix][n] = g(kpix, n),
//   A[m][k] = [k == m % 16]; D[m][n] must equal B[m % 16 + d][n].
// Test 3 (K-major A, 8-row groups at a stride that is NOT a multiple of 1024 B): the tile is 16 "image rows" of 10 pixel
//   slots (8 pixels + one halo slot each side), group g = the 8 pixels of image row g shifted by d - 1 slots:
//   start = base + d * 128, SBO = 1280; D[m][n] must equal A[(m / 8) * 10 + m % 8 + d][n].
// Test 4 (MN-major B with THREE 64-channel N atoms one row apart, LBO = 128 B): the three horizontal taps of a weight
//   gradient as one N = 192 MMA over a single staged row; D[m][64 a + c] must equal B[m % 16 + d + a][c].
// Each with the descriptor's base-offset field 0 and d & 7.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "msb_ptx.cuh"
using namespace msb;

__host__ __device__ inline int fa(int row, int k) { return (row * 3 + k * 7) % 200 + 1; }       // exact in bf16
__host__ __device__ inline int gb(int kpix, int n) { return (kpix * 5 + (n >> 3) * 23 + (n & 7)) % 200 + 1; }

__device__ inline void put(uint8_t* tile, int row, int col, float v) {      // 128-B-swizzled [row][64 bf16]
    const uint32_t off = row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(128, 1) probe(int test, int d, int use_base_offset, float* out /*[128][192]*/) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ta = smem;                 // 160 rows x 128 B = 20 KB
    uint8_t* tb = smem + 20 * 1024;     // 160 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 40 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    __syncthreads();
    if (test == 1 || test == 3) {
        for (int i = threadIdx.x; i < 160 * 64; i += blockDim.x) put(ta, i / 64, i % 64, (float)fa(i / 64, i % 64));
        for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) put(tb, i / 64, i % 64, (i % 64) == (i / 64) ? 1.f : 0.f);
    } else {
        for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) put(ta, i / 64, i % 64, (i % 64) == ((i / 64) % 16) ? 1.f : 0.f);
        for (int i = threadIdx.x; i < 160 * 64; i += blockDim.x) put(tb, i / 64, i % 64, (float)gb(i / 64, i % 64));
    }
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tbase, 512); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tbm = tbase;
    const uint32_t lane_addr = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    if (threadIdx.x == 0) {
        const uint64_t bo = use_base_offset ? ((uint64_t)(d & 7) << 49) : 0;
        if (test == 1 || test == 3) {
            const uint32_t idesc = ptx::make_idesc_bf16(128, 16, 0, 0);
            const uint64_t adesc = ptx::make_smem_desc_sw128(ptx::smem_u32(ta) + d * 128, 16, test == 3 ? 1280 : 1024) | bo;
            const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(tb), 16, 1024);
            ptx::umma_bf16(tbm, adesc, bdesc, idesc, 0u);
        } else if (test == 2) {
            const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 1);                       // B MN-major
            const uint64_t adesc = ptx::make_smem_desc_sw128(ptx::smem_u32(ta), 16, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(tb) + d * 128, 8192, 1024) | bo;
            ptx::umma_bf16(tbm, adesc, bdesc, idesc, 0u);
        } else {
            const uint32_t idesc = ptx::make_idesc_bf16(128, 192, 0, 1);                      // B MN-major, 3 atoms
            const uint64_t adesc = ptx::make_smem_desc_sw128(ptx::smem_u32(ta), 16, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(tb) + d * 128, 128, 1024) | bo;
            ptx::umma_bf16(tbm, adesc, bdesc, idesc, 0u);
        }
        ptx::umma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    for (int c = 0; c < 192; c += 16) {
        float v[16];
        ptx::tmem_ld16(tbm + lane_addr + c, v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 192 + c + j] = v[j];
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tbm, 512); }
}

int main() {
    float* dout; cudaMalloc(&dout, 128 * 192 * 4);
    static float h[128 * 192];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int test : {1, 2, 3, 4}) {
        printf("test %d (%s): mismatching elements of D per start offset d (rows of 128 B), base-offset field = 0 | d & 7\n", test,
               test == 1 ? "K-major A, M rows shifted" : (test == 2 ? "MN-major B, K rows shifted" : (test == 3 ? "K-major A, 8-row groups at SBO = 1280 B" : "MN-major B, N = 192 = three atoms at LBO = 128 B")));
        for (int d = 0; d <= (test >= 3 ? 2 : 9); ++d) {
            int bad[2] = {0, 0};
            for (int ubo = 0; ubo < 2; ++ubo) {
                cudaMemset(dout, 0, sizeof(h));
                probe<<<1, 128, 48 * 1024>>>(test, d, ubo, dout);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
                const int ncols = test == 2 ? 64 : (test == 4 ? 192 : 16);
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < ncols; ++n) {
                        const float want = test == 1 ? (float)fa(m + d, n) : (test == 2 ? (float)gb(m % 16 + d, n) :
                                           (test == 3 ? (float)fa((m / 8) * 10 + m % 8 + d, n) : (float)gb(m % 16 + d + n / 64, n % 64)));
                        if (h[m * 192 + n] != want) ++bad[ubo];
                    }
            }
            printf("  d = %d: %5d | %5d   of %d%s\n", d, bad[0], bad[1], 128 * (test == 2 ? 64 : (test == 4 ? 192 : 16)),
                   (bad[0] == 0 || bad[1] == 0) ? "" : "   (both wrong)");
        }
    }
    return 0;
}
