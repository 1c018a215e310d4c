// Microbenchmark 5: cost of tcgen05.commit + mbarrier hand-offs between batches of MMAs.
// One or two issuer warps; each runs  for (i..iters) { [wait bar[(i-S) % S]] ; issue M MMAs ; commit bar[i % S] }.
// wait = 0: commits only (is there a pipe bubble per commit?); wait = 1: the issuer also waits for the completion of
// the batch S iterations back (what a stage ring of depth S makes the real kernel do).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../poco_b200/csrc/common.cuh"
using namespace poco;

__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

template <int M>
__global__ void __launch_bounds__(128, 1) k(int N, int S, int wait, int nissue, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar[2][16];
    __shared__ unsigned long long done[2];
    __shared__ uint32_t tbase_s;
    __shared__ long long el[2];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) { mbar_init(smem_u32(&bar[threadIdx.x >> 4][threadIdx.x & 15]), 1); }
    if (threadIdx.x < 2) mbar_init(smem_u32(&done[threadIdx.x]), 1);
    mbar_fence_init();
    if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x00010001u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tbase_s;
    int cols = 32;
    while (cols < N) cols <<= 1;
    if (warp < nissue) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t a0 = smem_u32(smem) + 8192 + warp * 32768, b0 = smem_u32(smem) + 128 * 1024;
        const uint64_t da0 = umma_desc(a0, 6016, 128);
        const uint64_t db0 = umma_desc(b0, uint32_t(N) * 16u, 128);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int slot = i % S;
            if (wait && i >= S) mbar_wait(smem_u32(&bar[warp][slot]), uint32_t((i / S) - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tbase + uint32_t(warp * 256 + (i & 3) * cols);
#pragma unroll
                for (int t = 0; t < M; ++t)
                    umma_ss(d, da0 + uint64_t((t % 9) * 37), db0 + uint64_t((t % 18) * 2 * N), idesc, t ? 1u : 0u);
                umma_commit(smem_u32(&bar[warp][slot]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&done[warp]));
        __syncwarp();
        mbar_wait(smem_u32(&done[warp]), 0);
        if (elect_one()) el[warp] = clock64() - t0;
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = nissue == 2 ? max(el[0], el[1]) : el[0];
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int M>
void run(int N, int S, int wait, int nissue, long long* d) {
    const int iters = 400;
    cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k<M><<<148, 128, 220 * 1024>>>(N, S, wait, nissue, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d,%d,%d,%d,%d,%.1f\n", N, M, S, wait, nissue, double(h) / (double(iters) * M * nissue));
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    printf("N,mmas_per_commit,ring,wait,issuers,cycles_per_mma\n");
    for (int N : {32, 64, 128})
        for (int nissue : {1, 2})
            for (int wait : {0, 1})
                for (int S : {1, 2, 4, 8}) {
                    if (!wait && S != 2) continue;
                    run<9>(N, S, wait, nissue, d); run<18>(N, S, wait, nissue, d); run<36>(N, S, wait, nissue, d); run<72>(N, S, wait, nissue, d);
                }
    return 0;
}
