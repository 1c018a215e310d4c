// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16) with both operands in shared
// memory in the no-swizzle K-major layout, as a function of N -- the hardware floor the conv kernel
// is designed against.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../poco_b200/csrc/common.cuh"
using namespace poco;

__global__ void __launch_bounds__(128, 1) k(int N, int M, int iters, int distinct, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 512);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_f16(M, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t aoff = distinct ? uint32_t(i % 16) * 4096u + uint32_t(i % 7) * 16u : 0u;
                const uint64_t da = umma_desc(a0 + aoff, 4096, 128);
                const uint64_t db = umma_desc(b0 + (distinct ? uint32_t(i % 4) * 8192u : 0u), uint32_t(N) * 16u, 128);
                umma_f16(tbase, da, db, idesc, i > 0);
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if (elect_one()) { t1 = clock64(); if (blockIdx.x == 0) { out[0] = t1 - t0; } }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("grid,M,N,distinct,cycles_per_mma\n");
    for (int grid : {1, 148})
        for (int M : {128, 64})
            for (int N : {16, 32, 64, 128, 256})
                for (int distinct : {0, 1}) {
                    long long h = 0;
                    for (int rep = 0; rep < 2; ++rep) {
                        k<<<grid, 128, 200 * 1024>>>(N, M, iters, distinct, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    printf("%d,%d,%d,%d,%.1f\n", grid, M, N, distinct, double(h) / iters);
                }
    return 0;
}
