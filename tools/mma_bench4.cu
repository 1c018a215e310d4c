// Microbenchmark 4: does the A-operand start alignment / tap pattern / LBO change the tensor-pipe cost of an SS-mode
// M=128 K=16 tcgen05.mma?  One issuing thread, fully unrolled loop of 18 MMAs (9 taps x 2 K steps) per "tile",
// accumulator switches every tile.  Patterns: 0 = all taps at the same aligned address, 1 = the conv's 9 taps
// ((r-1)*Wp + (s-1) pixels, Wp = 58: 16-byte granular starts), 2 = dy taps only aligned (Wp = 64, dx = 0),
// 3 = taps shifted by multiples of 8 pixels (128-byte aligned starts), 4 = dx = +-1 with Wp = 64.
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

template <int PATTERN>
__global__ void __launch_bounds__(128, 1) k(int N, int lbo, int base_shift, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1);
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
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        // A region: 4 planes of `lbo` bytes starting at 8 KB (room for negative tap shifts); B region at 128 KB
        const uint32_t a0 = smem_u32(smem) + 8192 + base_shift * 16, b0 = smem_u32(smem) + 128 * 1024;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            const uint64_t da0 = umma_desc(a0, lbo, 128);
            const uint64_t db0 = umma_desc(b0, uint32_t(N) * 16u, 128);
            t0 = clock64();
#pragma unroll 1
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tbase + uint32_t((i & 3) * cols);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    int sh;
                    if (PATTERN == 0) sh = 0;
                    else if (PATTERN == 1) sh = (t / 3 - 1) * 58 + (t % 3 - 1);
                    else if (PATTERN == 2) sh = (t / 3 - 1) * 64;
                    else if (PATTERN == 3) sh = (t - 4) * 8;
                    else sh = (t / 3 - 1) * 64 + (t % 3 - 1);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        umma_ss(d, da0 + uint64_t(int64_t(sh)) + uint64_t(ks * 2 * (lbo >> 4)),
                                db0 + uint64_t((t * 4 + ks * 2) * N), idesc, (t | ks) ? 1u : 0u);
                }
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if (elect_one()) { t1 = clock64(); if (blockIdx.x == 0) out[0] = t1 - t0; }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int PATTERN>
void run(int N, int lbo, int base_shift, long long* d) {
    const int iters = 400;
    cudaFuncSetAttribute(k<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k<PATTERN><<<148, 128, 220 * 1024>>>(N, lbo, base_shift, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d,%d,%d,%d,%.1f\n", PATTERN, N, lbo, base_shift, double(h) / (iters * 18));
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    printf("pattern,N,lbo,base_shift_px,cycles_per_mma\n");
    for (int N : {32, 64, 128, 256})
        for (int lbo : {4096, 6016, 6144}) {
            if (9 * 4 * N * 16 > 90 * 1024) continue;
            for (int bs : {0, 3}) {
                run<0>(N, lbo, bs, d); run<1>(N, lbo, bs, d); run<2>(N, lbo, bs, d); run<3>(N, lbo, bs, d); run<4>(N, lbo, bs, d);
            }
        }
    return 0;
}
