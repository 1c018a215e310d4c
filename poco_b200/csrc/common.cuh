// poco_b200 -- shared device helpers (sm_100a only): mbarrier, bulk-copy (TMA engine, UBLKCP),
// cp.async, tcgen05 (UMMA / TMEM) wrappers and the activation-layout index math.
//
// Activation layout ("planar-8 padded", fp16):   [C/8][N][H+2][W+2][8]
//   * one *plane* holds 8 consecutive channels of every pixel of every crop (16 B per pixel);
//   * every crop carries a 1-pixel zero halo that no kernel ever writes, so a 3x3/pad-1 window of
//     output pixel q (q = padded-linear index inside the plane) is the 9 linear shifts
//     q + (r-1)*(W+2) + (s-1): an implicit-GEMM A-tile is a *contiguous* 1-D run of the plane and is
//     fetched with ONE bulk copy per plane -- no im2col, no tensor map;
//   * 16 B per pixel is exactly the 8x16B core-matrix row of the UMMA no-swizzle K-major canonical
//     layout, so the landed bytes are MMA-ready (SBO = 128 B, LBO = plane pitch in shared memory).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if !defined(__CUDA_ARCH_FEAT_SM100_ALL) && defined(__CUDA_ARCH__)
#error "poco_b200 kernels are written for sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif

namespace poco {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// programmatic dependent launch (no-ops when the kernel was launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one lane of a fully converged warp (PTX elect.sync); lets ptxas keep the guarded operands uniform
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Watchdog for spin waits: a protocol bug must surface as a launch failure with a location, never as
// a hung GPU.  Cold path only (the clock is read every 1024 failed polls).
static __device__ __noinline__ void spin_timeout(int tag, unsigned long long t0) {
    if (global_timer_ns() - t0 > 2000000000ull) {
        printf("poco_b200: wait timed out (tag %d) block (%d,%d) thread %d\n", tag, blockIdx.x, blockIdx.y, threadIdx.x);
        __trap();
    }
}
__device__ __forceinline__ void mbar_wait_tag(uint32_t bar, uint32_t parity, int tag) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0u) {
            if (t0 == 0) t0 = global_timer_ns();
            spin_timeout(tag, t0);
        }
    }
}

// Wait of a role that shares its SM sub-partition with busy warps (producer / MMA-issuer warps next to epilogue
// warps): try_wait with a suspend-time hint parks the warp in hardware until the phase completes (or the hint expires)
// instead of spinning through the issue slots of its neighbours.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!ok) {
        if ((++spins & 1023u) == 0u) {          // (an iteration parks for up to 20 us: this is seconds of waiting)
            if (t0 == 0) t0 = global_timer_ns();
            spin_timeout(-1, t0);
        }
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// async copies
// ---------------------------------------------------------------------------------------------
// 1-D bulk copy global -> shared, executed by the TMA engine (SASS: UBLKCP), completes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N committed bulk-store groups still have to READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// 16-byte cp.async (LDGSTS) with zero-fill when !valid; .ca: neighbouring filter taps of a stride-2
// gather hit the same 32-byte sectors, let L1 absorb the re-reads
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier is arrived-on once all previously issued MMAs of this thread retire
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate, M=128, K=16 per instruction
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor, no-swizzle K-major canonical layout (cute::UMMA::SmemDescriptor):
//   core matrix = 8 rows x 16 B (rows 16 B apart); SBO = byte pitch between 8-row groups (M/N dir);
//   LBO = byte pitch between the two 16-byte K halves of one K=16 MMA.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);            // [0,14)  start address >> 4
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;   // [16,30) leading-dim byte offset >> 4
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;   // [32,46) stride-dim byte offset >> 4
    d |= 1ull << 46;                                                // [46,48) descriptor version (sm_100)
    // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
    return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): fp16 x fp16 -> fp32, K-major A and B
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)            // c_format = F32
           | (0u << 7)          // a_format = F16
           | (0u << 10)         // b_format = F16
           | (0u << 15)         // a_major = K
           | (0u << 16)         // b_major = K
           | ((N >> 3) << 17)   // n_dim
           | ((M >> 4) << 24);  // m_dim
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns (one row per thread)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
}

// ---------------------------------------------------------------------------------------------
// MMA issue helpers shared by conv_tc.cu and bblock_tc.cu
// ---------------------------------------------------------------------------------------------
// -DPOCO_MBAR_WATCHDOG: every mbarrier wait gets a 2 s watchdog that reports its source line and traps
// (bring-up builds only: the bookkeeping in the hot wait loops costs 8-14 % of every conv, measured).
// The global-memory flag spin of the chain protocol always carries the watchdog.
#ifdef POCO_MBAR_WATCHDOG
#define MBAR_WAIT(bar, parity) mbar_wait_tag(bar, parity, __LINE__)
#else
#define MBAR_WAIT(bar, parity) mbar_wait(bar, parity)
#endif

// Issue every MMA of one shared-memory stage as straight-line code.  The issuing thread is the
// bottleneck for small N: measured (profiles/r01_mma_issue_rate.csv) the tensor pipe accepts an SS-mode
// M=128 K=16 MMA every max(N/2, 32 + N/4) cycles (operand fetch at 128 B/clk), but a descriptor that is
// built in vector registers and moved with R2UR costs the thread ~77 cycles per MMA.  Here every
// descriptor is `base + compile-time multiple of a uniform stride`, so ptxas keeps the chain in uniform
// registers: ~3 uniform ALU ops per UTCHMMA.
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return (uint64_t(hi) << 32) | lo; }

// POCO_ISSUE_BATCH (round 2): UTCHMMA reads its descriptors from uniform registers and holds them until the tensor pipe
// dequeues the instruction; a uniform-datapath op that overwrites one of them stalls (short scoreboard) until then.  The
// lean "3 uniform ops per MMA" chain reused a handful of registers every 2-3 MMAs, so only ~3 MMAs were ever queued and the
// pipe ran at 60-72 cycles per MMA instead of its 40-64 (ncu source page: 80 % of the issuer's samples were short_sb on
// UIADD3; tools/mma_bench7.cu reaches the hardware rate with the same operand geometry).  Here the descriptors of a batch
// of MMAs are materialised in ordinary registers first; ptxas then moves each into its own uniform register pair right
// before its UTCHMMA, so a whole batch can sit in the queue.  0 = the old uniform chain.
#ifndef POCO_ISSUE_BATCH
#define POCO_ISSUE_BATCH 12
#endif
template <int TAPS, int KS>
__device__ __forceinline__ void issue_linear(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, const uint32_t (&sh)[9], uint32_t a_kstep,
                                             uint32_t b_kstep, uint32_t b_tap, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t acc0) {
#if POCO_ISSUE_BATCH > 0
    constexpr int TOTAL = TAPS * KS;
    constexpr int BATCH = POCO_ISSUE_BATCH < TOTAL ? POCO_ISSUE_BATCH : TOTAL;
#pragma unroll
    for (int j0 = 0; j0 < TOTAL; j0 += BATCH) {
        uint32_t al[BATCH], bl[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int j = j0 + i, t = j / KS, k = j % KS;
            if (j < TOTAL) {
                al[i] = a_lo + sh[t] + uint32_t(k) * a_kstep;       // sh: where tap t starts inside the landed run
                bl[i] = b_lo + uint32_t(t) * b_tap + uint32_t(k) * b_kstep;
            }
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i)
            if (j0 + i < TOTAL) asm volatile("" : "+r"(al[i]), "+r"(bl[i]));        // all of the batch live in registers here
#pragma unroll
        for (int i = 0; i < BATCH; ++i)
            if (j0 + i < TOTAL) umma_f16(d_tmem, desc64(desc_hi, al[i]), desc64(desc_hi, bl[i]), idesc, (j0 + i) ? 1u : acc0);
    }
#else
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
        const uint32_t at = a_lo + sh[t];
        const uint32_t bt = b_lo + uint32_t(t) * b_tap;
#pragma unroll
        for (int k = 0; k < KS; ++k)
            umma_f16(d_tmem, desc64(desc_hi, at + uint32_t(k) * a_kstep), desc64(desc_hi, bt + uint32_t(k) * b_kstep), idesc,
                     (t | k) ? 1u : acc0);
    }
#endif
}

}  // namespace poco
