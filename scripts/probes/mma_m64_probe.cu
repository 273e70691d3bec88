// Probe: which TMEM lanes a tcgen05.mma with M = 64 (cta_group::1, A from TMEM) reads its A rows from and writes its D rows to.
//   A[lane][k] = lane + 1 for k = 0, else 0 (all 128 lanes x 8 columns written with tcgen05.st), B = all ones,
//   D pre-filled with -1  =>  after the MMA, D[lane L][n] = (A lane used by the row that lives in D lane L) + 1, or -1 if untouched.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "msb_ptx.cuh"
using namespace msb;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int M, int accumulate_first, float* out /*[128][64]*/) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3f803f80u;      // bf16 1.0 pairs
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tbase, 512); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tb = tbase;
    const uint32_t lane_addr = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    {
        uint32_t r[16];
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(-1.0f);
        for (int c = 0; c < 64; c += 16) st16(tb + lane_addr + c, r);               // D = -1
        const __nv_bfloat16 v = __float2bfloat16_rn((float)(threadIdx.x + 1));
        for (int j = 0; j < 16; ++j) r[j] = 0;
        r[0] = (uint32_t)__bfloat16_as_ushort(v);                                    // k = 0 in the low half of column 0
        st16(tb + lane_addr + 256, r);                                              // A slab at column 256
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::make_idesc_bf16(M, 64, 0, 0);
        const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem), 16, 1024);
        umma_ts(tb, tb + 256, bdesc, idesc, accumulate_first);
        ptx::umma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    for (int c = 0; c < 64; c += 16) {
        float v[16];
        ptx::tmem_ld16(tb + lane_addr + c, v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 64 + c + j] = v[j];
    }
    ptx::tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 512); }
}

int main() {
    float* dout; cudaMalloc(&dout, 128 * 64 * 4);
    static float h[128 * 64];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int M : {128, 64}) for (int accf : {0, 1}) {
        probe<<<1, 128, 16384>>>(M, accf, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d: CUDA error: %s\n", M, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M = %d accumulate=%d: D lane -> value in columns 0 / 31 / 32 / 63 (A lane + 1; -1 = untouched)\n", M, accf);
        for (int L = 0; L < 128; ++L) {
            printf(" L%3d:%4.0f %4.0f %4.0f %4.0f%s", L, h[L * 64], h[L * 64 + 31], h[L * 64 + 32], h[L * 64 + 63], (L % 4 == 3) ? "\n" : " |");
        }
    }
    return 0;
}
