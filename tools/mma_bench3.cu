// Microbenchmark 3: hardware rate of small-N tcgen05.mma with a lean, fully unrolled issue loop
// (descriptors advance by compile-time constants), NACC rotating accumulators, 1/2/4 issuer warps.
// Separates "the issuing thread is slow" from "the tensor pipe is slow".
#include <cstdio>
#include <cuda_runtime.h>
#include "../poco_b200/csrc/common.cuh"
using namespace poco;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc)
        : "memory");
}

template <int NACC, int MODE>
__global__ void __launch_bounds__(128, 1) k(int N, int M, int iters, int nissue, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar[4];
    __shared__ uint32_t tbase_s;
    __shared__ long long el[4];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) mbar_init(smem_u32(&bar[threadIdx.x]), 1);
    mbar_fence_init();
    if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x00010001u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tbase_s;
    int cols = 32;
    while (cols < N) cols <<= 1;
    if (warp < nissue) {
        const uint32_t idesc = umma_idesc_f16(M, N);
        const uint32_t a0 = smem_u32(smem) + warp * 20480, b0 = smem_u32(smem) + 96 * 1024 + warp * 8192;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            const uint64_t da0 = umma_desc(a0, 4096, 128);
            const uint64_t db0 = umma_desc(b0, uint32_t(N) * 16u, 128);
            const uint32_t d0 = tbase + uint32_t(warp * NACC * cols);
            t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t d = d0 + uint32_t((u % NACC) * cols);
                    if (MODE == 0) umma_ss(d, da0 + uint64_t(u * 37), db0 + uint64_t((u & 3) * 16), idesc);   // A shifts by 37 pixels (592 B), B by 256 B
                    else umma_ts(d, tbase + 480u + uint32_t(u & 3) * 8u, db0 + uint64_t((u & 3) * 16), idesc);
                }
            }
            umma_commit(smem_u32(&bar[warp]));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar[warp]), 0);
        if (elect_one()) { t1 = clock64(); el[warp] = t1 - t0; }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long m = 0;
        for (int i = 0; i < nissue; ++i) m = max(m, el[i]);
        out[0] = m;
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int NACC, int MODE>
void run(int N, int nissue, long long* d) {
    const int iters = 4000;
    int cols = 32; while (cols < N) cols <<= 1;
    if (nissue * NACC * cols > 448) return;
    cudaFuncSetAttribute(k<NACC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k<NACC, MODE><<<148, 128, 200 * 1024>>>(N, 128, iters, nissue, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d,%d,%d,%d,%.1f\n", MODE, N, NACC, nissue, double(h) / (iters * nissue));
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    printf("mode,N,nacc,nissue,cycles_per_mma\n");
    for (int N : {16, 32, 64, 96, 128, 256})
        for (int nissue : {1, 2, 4}) {
            run<1, 0>(N, nissue, d); run<2, 0>(N, nissue, d); run<4, 0>(N, nissue, d);
            run<1, 1>(N, nissue, d); run<4, 1>(N, nissue, d);
        }
    return 0;
}
