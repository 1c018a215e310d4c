// Microbenchmark 6: what slows the tensor pipe's shared-memory operand fetch?  Two issuer warps (commit bubbles hidden,
// bench 5) run M=128 K=16 SS MMAs while `nbg` background warps do one of:
//   bg = 0 nothing; 1 mbarrier.try_wait spin on a barrier that never completes; 2 ld.shared.v4 stream (conflict-free);
//   3 bulk copies global -> shared (TMA writes) issued by one lane per background warp, 4 KB each, back to back.
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

__global__ void __launch_bounds__(384, 1) k(int N, int bg, int nbg, int iters, const uint8_t* gsrc, long long* out, int* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar[2][4];
    __shared__ unsigned long long done[2], never, cpbar[16];
    __shared__ uint32_t tbase_s;
    __shared__ volatile int stop;
    __shared__ long long el[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 8) mbar_init(smem_u32(&bar[threadIdx.x >> 2][threadIdx.x & 3]), 1);
    if (threadIdx.x < 2) mbar_init(smem_u32(&done[threadIdx.x]), 1);
    if (threadIdx.x == 2) mbar_init(smem_u32(&never), 1);
    if (threadIdx.x >= 16 && threadIdx.x < 32) mbar_init(smem_u32(&cpbar[threadIdx.x - 16]), 1);
    if (threadIdx.x == 0) stop = 0;
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
    if (warp < 2) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t a0 = smem_u32(smem) + 8192 + warp * 32768, b0 = smem_u32(smem) + 96 * 1024;
        const uint64_t da0 = umma_desc(a0, 6016, 128);
        const uint64_t db0 = umma_desc(b0, uint32_t(N) * 16u, 128);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int slot = i & 3;
            if (i >= 4) mbar_wait(smem_u32(&bar[warp][slot]), uint32_t((i >> 2) - 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tbase + uint32_t(warp * 256 + (i & 1) * cols);
#pragma unroll
                for (int t = 0; t < 18; ++t)
                    umma_ss(d, da0 + uint64_t((t % 9) * 37), db0 + uint64_t(t * 2 * N), idesc, t ? 1u : 0u);
                umma_commit(smem_u32(&bar[warp][slot]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&done[warp]));
        __syncwarp();
        mbar_wait(smem_u32(&done[warp]), 0);
        if (elect_one()) el[warp] = clock64() - t0;
        __syncwarp();
        if (warp == 0 && lane == 0) {
            mbar_wait(smem_u32(&done[1]), 0);
            stop = 1;
            mbar_arrive(smem_u32(&never));
        }
    } else if (warp - 2 < nbg) {
        const int w = warp - 2;
        if (bg == 1) {
            mbar_wait(smem_u32(&never), 0);
        } else if (bg == 2) {
            int acc = 0;
            const uint32_t base = smem_u32(smem) + 160 * 1024 + w * 2048 + lane * 16;
            while (!stop) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int4 v;
                    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + r * 512));
                    acc += v.x + v.y + v.z + v.w;
                }
            }
            if (acc == 12345) sink[0] = acc;
        } else if (bg == 3) {
            uint32_t ph = 0;
            const uint32_t dst = smem_u32(smem) + 160 * 1024 + w * 4096;
            const uint8_t* src = gsrc + (size_t(blockIdx.x) * 16 + w) * 65536;
            int off = 0;
            while (!stop) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(smem_u32(&cpbar[w]), 4096);
                    bulk_g2s(dst, src + off, 4096, smem_u32(&cpbar[w]));
                }
                __syncwarp();
                mbar_wait(smem_u32(&cpbar[w]), ph);
                ph ^= 1u;
                off = (off + 4096) & 65535;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = max(el[0], el[1]);
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    int* sink; cudaMalloc(&sink, 64);
    uint8_t* g; cudaMalloc(&g, size_t(148) * 16 * 65536); cudaMemset(g, 1, size_t(148) * 16 * 65536);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    printf("N,bg,nbg,cycles_per_mma\n");
    const int iters = 400;
    for (int N : {32, 64, 128})
        for (int bg : {0, 1, 2, 3})
            for (int nbg : {1, 4, 9}) {
                if (bg == 0 && nbg != 1) continue;
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    k<<<148, 384, 220 * 1024>>>(N, bg, nbg, iters, g, d, sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("%d,%d,%d,%.1f\n", N, bg, nbg, double(h) / (double(iters) * 18 * 2));
            }
    return 0;
}
