// Microbenchmark 2: is the ~62-77 cycle floor of a small-N tcgen05.mma (SS mode) a property of the
// instruction, or of the accumulator dependency chain / the single issuing thread?  Issues MMAs that
// rotate over `nacc` independent TMEM accumulators from `nissue` issuer warps and reports the CTA-level
// cycles per MMA.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench2 tools/mma_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../poco_b200/csrc/common.cuh"
using namespace poco;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: SS, mode 1: TS (A operand in TMEM columns 480..511)
__global__ void __launch_bounds__(128, 1) k(int N, int M, int iters, int nacc, int nissue, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar[2];
    __shared__ uint32_t tbase_s;
    __shared__ long long el[2];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tbase_s;
    int cols = 32;
    while (cols < N) cols <<= 1;
    if (warp < nissue) {
        const uint32_t idesc = umma_idesc_f16(M, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t aoff = uint32_t(i % 16) * 4096u + uint32_t(i % 7) * 16u;
                const uint64_t da = umma_desc(a0 + aoff, 4096, 128);
                const uint64_t db = umma_desc(b0 + uint32_t(i % 4) * 8192u, uint32_t(N) * 16u, 128);
                const uint32_t d = tbase + uint32_t((warp * nacc + i % nacc) * cols);
                if (mode == 0) umma_f16(d, da, db, idesc, i >= nacc);
                else umma_f16_ts(d, tbase + 480u + uint32_t(i % 4) * 8u, db, idesc, i >= nacc);
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
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = nissue == 2 ? max(el[0], el[1]) : el[0];
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("mode,grid,M,N,nacc,nissue,cycles_per_mma\n");
    for (int mode : {0, 1})
        for (int grid : {148})
            for (int M : {128})
                for (int N : {16, 32, 64, 128, 256})
                    for (int nissue : {1, 2})
                        for (int nacc : {1, 2, 4, 8}) {
                            int cols = 32; while (cols < N) cols <<= 1;
                            if (nissue * nacc * cols > 448) continue;
                            long long h = 0;
                            for (int rep = 0; rep < 2; ++rep) {
                                k<<<grid, 128, 200 * 1024>>>(N, M, iters, nacc, nissue, mode, d);
                                cudaError_t e = cudaDeviceSynchronize();
                                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                            }
                            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                            printf("%d,%d,%d,%d,%d,%d,%.1f\n", mode, grid, M, N, nacc, nissue, double(h) / (iters * nissue));
                        }
    return 0;
}
