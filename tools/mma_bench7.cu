// Microbenchmark 7: the real conv's operand geometry.  Two issuer warps alternate "tiles"; a tile = TAPS x KS MMAs with
// A = base + tap shift ((r-1)*Wp + (s-1) pixels) + k * 2 * lbo, B = w + tap * (cin8 * N * 16) + k * (2 * N * 16),
// accumulators rotate over `nacc` TMEM buffers.  Variants isolate which parameter moves the per-MMA cost.
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

template <int KS>
__global__ void __launch_bounds__(128, 1) k(int N, int lbo, int Wp, int nacc, int b_mode, int a_mode, int iters, long long* out, int data_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar[2][4];
    __shared__ unsigned long long done[2];
    __shared__ uint32_t tbase_s;
    __shared__ long long el[2];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) mbar_init(smem_u32(&bar[threadIdx.x >> 2][threadIdx.x & 3]), 1);
    if (threadIdx.x < 2) mbar_init(smem_u32(&done[threadIdx.x]), 1);
    mbar_fence_init();
    if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
    for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t(i) * 2654435761u) ^ (blockIdx.x * 40503u);
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        // data_mode 0: constant ~1.0; 1: random fp16 in [-2, 2) (random sign / exponent 0x38..0x3f / mantissa)
        reinterpret_cast<uint32_t*>(smem)[i] = data_mode == 0 ? 0x3c003c00u + i % 7 : ((h & 0x87ff87ffu) | 0x38003800u) ^ ((h >> 3) & 0x04000400u);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tbase_s;
    int cols = 32;
    while (cols < N) cols <<= 1;
    if (warp < 2) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        // A: two stage buffers of 2*KS planes each (one per issuer), 2 KB of slack in front for negative shifts
        const uint32_t a0 = smem_u32(smem) + 2048 + warp * (2 * KS * lbo + 4096), b0 = smem_u32(smem) + 120 * 1024;
        const uint64_t da0 = umma_desc(a0, lbo, 128);
        const uint64_t db0 = umma_desc(b0, uint32_t(N) * 16u, 128);
        const uint32_t b_tap = b_mode == 0 ? uint32_t(2 * KS * N) : 0u, b_k = b_mode == 0 ? uint32_t(2 * N) : 0u;   // descriptor units
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int slot = i & 3;
            if (i >= 4) mbar_wait(smem_u32(&bar[warp][slot]), uint32_t((i >> 2) - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tbase + uint32_t(((2 * i + warp) % nacc) * cols);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int sh = a_mode == 0 ? (t / 3 - 1) * Wp + (t % 3 - 1) : 0;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                        umma_ss(d, da0 + uint64_t(int64_t(sh)) + uint64_t(ks * 2 * (lbo >> 4)),
                                db0 + uint64_t(t * b_tap + ks * b_k), idesc, (t | ks) ? 1u : 0u);
                }
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
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = max(el[0], el[1]);
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

static int g_data_mode = 0;
template <int KS>
void run(const char* tag, int N, int lbo, int Wp, int nacc, int b_mode, int a_mode, long long* d) {
    const int iters = 300;
    cudaFuncSetAttribute(k<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k<KS><<<148, 128, 225 * 1024>>>(N, lbo, Wp, nacc, b_mode, a_mode, iters, d, g_data_mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s,%d,%d,%d,%d,%d,%d,%d,%.1f\n", tag, N, KS, lbo, Wp, nacc, b_mode, a_mode, double(h) / (double(iters) * 9 * KS * 2));
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    printf("tag,N,ksteps,lbo,Wp,nacc,b_mode,a_mode,cycles_per_mma\n");
    run<4>("conv64_real", 64, 3072, 30, 8, 0, 0, d);       // 64->64 @28: 73 KB of weights walked, 8 accumulators
    run<4>("conv64_nacc2", 64, 3072, 30, 2, 0, 0, d);
    run<4>("conv64_sameB", 64, 3072, 30, 8, 1, 0, d);      // every MMA reads the same 2 KB of B
    run<4>("conv64_sameA", 64, 3072, 30, 8, 0, 1, d);      // no tap shifts
    run<4>("conv64_lbo6144", 64, 6144, 30, 8, 0, 0, d);
    run<2>("conv32_real", 32, 6016, 58, 8, 0, 0, d);       // 32->32 @56 (one tile of a G = 2 run)
    run<2>("conv32_sameB", 32, 6016, 58, 8, 1, 0, d);
    g_data_mode = 1;
    run<4>("conv64_real_randomdata", 64, 3072, 30, 8, 0, 0, d);
    run<2>("conv32_real_randomdata", 32, 6016, 58, 8, 0, 0, d);
    return 0;
}
