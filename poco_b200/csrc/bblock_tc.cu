// poco_b200 -- one residual BasicBlock as ONE tcgen05 launch (sm_100a):
//     out = ReLU(BN2(conv2(ReLU(BN1(conv1(in))))) + in)          (hrnet.py:42-58, downsample is None)
// for the 32-channel, 56x56 branch of the HRNet modules (hrnet.py:140-186): 64 of the 311 conv launches of a
// POCO-CLIFF / HRNet-W32 forward.  As two poco_conv launches a block pays the fixed launch gap + pipeline fill twice,
// writes the intermediate tensor to HBM, reads it back with a second halo, and reads the block input a second time as
// the residual (257 MB of HBM traffic per block at batch 256).  Here the intermediate never leaves the SM:
//
//   work unit u = G = 3 conv2 tiles (384 consecutive padded-linear output pixels [p0, p0 + 384)).  conv2 reaches
//   R = W + 3 pixels to either side, so the unit needs conv1's output on [p0 - R, p0 + 384 + R): it computes
//   T1 = G + 1 = 4 conv1 tiles starting at p0 - 64 (halo recompute: 7 tiles of MMAs per 3 tiles of output, 1.17x)
//   from ONE bulk copy per input plane of 512 + 2R pixels.  Epilogue 1 turns the four TMEM accumulators into
//   fp16 ReLU(acc + shift1), zero at halo / out-of-range pixels (that IS conv2's zero padding), and stores them into
//   shared memory in the planar operand layout ([plane][pixel][8 ch], 16 B per pixel = one core-matrix row), so conv2's
//   nine taps are nine shifted descriptors into those bytes exactly like conv1's into the landed input run.
//   Epilogue 2 adds shift2 and the residual (the block input, re-read through L2 right after the bulk copy fetched
//   it) and writes the 16-byte pixels of the four output planes.
//
// HBM traffic per block: input once + output once (103 MB at batch 256); the kernel is bound by the tensor pipe at
// its small-N rate (an M = 128, N = 32, K = 16 MMA every 40 cycles: 7 x 18 x 40 = 5040 cycles per unit).
//
// Roles (352 threads, one persistent CTA per SM, units dealt round-robin): warp 0 producer (weights once, then the
// input runs, double buffered), warp 1 issues conv1's MMAs and warp 2 conv2's, each unit by unit behind its own
// barriers, so that the tensor pipe always has the other conv's tiles queued while an epilogue runs; warps 3-10 epilogue
// (two sets of four TMEM lane groups; a set takes alternate tiles).  TMEM: 2 x (4 + 3) accumulators of 32 columns.
// mbarriers, each a two-deep ring over the local unit index j (buffer j & 1, use j >> 1):
//   in_full / in_free (producer <-> C1), acc1_full / acc1_free (C1 <-> epilogue 1), mid_full / mid_free
//   (epilogue 1 <-> C2), acc2_full / acc2_free (C2 <-> epilogue 2).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "internal.h"

namespace poco {

namespace {

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

constexpr int kC = 32;                  // channels (= N of every MMA)
constexpr int kPlanes = kC / 8;
constexpr int kG = 3;                   // conv2 tiles per unit
constexpr int kT1 = kG + 1;             // conv1 tiles per unit
constexpr int kTile = 128;
constexpr int kLead = 64;               // conv1 starts this many pixels before the unit's first output pixel (>= W + 3)
constexpr int kSlab = kC * 16;          // one (tap, 8-channel) weight slab: 32 output channels x 16 B
constexpr int kWBytes = 9 * kPlanes * kSlab;        // one conv's weights: 18 KB
constexpr int kMidPitch = kT1 * kTile * 16;         // one plane of the intermediate: 8 KB
constexpr int kHeader = 1024;
constexpr int kThreads = 352;          // producer, two issuers, eight epilogue warps
constexpr int kTmemCols = 512;          // 2 x (4 + 3) x 32 = 448 used

struct BlockParams {
    const __half* in;
    __half* out;
    const __half* w1;
    const __half* w2;
    const float* b1;
    const float* b2;
    long long in_plane, out_plane;      // plane strides in pixels
    int H, W;
    int P;                              // N * (H + 2) * (W + 2)
    int num_units;
    int in_pitch;                       // bytes between the planes of an input run in shared memory
    int run_bytes;                      // bytes of one input run: (4 * 128 + 2 R) * 16
    unsigned long long* prof;           // bring-up (POCO_BBLOCK_PROF): per issuer [total, in_full, acc1_free, mid_full, acc2_free, issue, units, ctas]
};

struct Header {
    unsigned long long in_full[2], in_free[2], acc1_full[2], acc1_free[2], mid_full[2], mid_free[2], acc2_full[2], acc2_free[2];
    unsigned long long w_full;
    uint32_t tmem_base;
    uint32_t pad_;
    float bias[2][kC];
};
static_assert(sizeof(Header) <= kHeader, "header too large");

__global__ void __launch_bounds__(kThreads, 1) basic_block_kernel(const BlockParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Header* hdr = reinterpret_cast<Header*>(smem);
    uint8_t* w_smem = smem + kHeader;                               // [conv][tap][plane][32][8]
    uint8_t* in_smem = w_smem + 2 * kWBytes;                        // [2][plane][run]
    const int in_buf_bytes = kPlanes * p.in_pitch;
    uint8_t* mid_smem = in_smem + 2 * in_buf_bytes;                 // [2][plane][512 pixels]
    constexpr int mid_buf_bytes = kPlanes * kMidPitch;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int Wp = p.W + 2, HpWp = (p.H + 2) * Wp, R = Wp + 1;
    const int my_units = (p.num_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

    if (threadIdx.x < 16) {
        unsigned long long* bars = hdr->in_full;                    // the sixteen ring barriers are contiguous
        const int which = threadIdx.x >> 1;                         // 0 in_full 1 in_free 2 acc1_full 3 acc1_free 4 mid_full 5 mid_free 6 acc2_full 7 acc2_free
        const uint32_t count = (which == 3 || which == 4 || which == 7) ? 8u : 1u;      // the eight epilogue warps / one commit or producer
        mbar_init(smem_u32(bars + threadIdx.x), count);
    }
    if (threadIdx.x == 16) mbar_init(smem_u32(&hdr->w_full), 1);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * kC) {
        const int i = threadIdx.x - 64;
        hdr->bias[i / kC][i % kC] = (i < kC ? p.b1 : p.b2)[i % kC];
    }
    mbar_fence_init();
    if (warp == 1) tmem_alloc(smem_u32(&hdr->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    auto acc1_col = [&](uint32_t b, int t) { return (b * kT1 + uint32_t(t)) * uint32_t(kC); };
    auto acc2_col = [&](uint32_t b, int g) { return (2u * kT1 + b * kG + uint32_t(g)) * uint32_t(kC); };

    if (warp == 0) {
        // ============================================================ producer
        if (elect_one()) {
            const uint32_t bar = smem_u32(&hdr->w_full);
            mbar_arrive_expect_tx(bar, 2u * kWBytes);
            bulk_g2s(smem_u32(w_smem), p.w1, kWBytes, bar);
            bulk_g2s(smem_u32(w_smem) + kWBytes, p.w2, kWBytes, bar);
        }
        __syncwarp();
        for (int j = 0; j < my_units; ++j) {
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            MBAR_WAIT(smem_u32(&hdr->in_free[b]), par ^ 1u);        // C1 of the unit two back has retired
            if (elect_one()) {
                const uint32_t bar = smem_u32(&hdr->in_full[b]);
                mbar_arrive_expect_tx(bar, uint32_t(kPlanes) * uint32_t(p.run_bytes));
                const long long q0 = unit * (kG * kTile) - kLead - R;       // (>= -8 KB guard, see the launcher)
                const __half* src = p.in + q0 * 8;
                const uint32_t dst = smem_u32(in_smem) + b * uint32_t(in_buf_bytes);
                for (int pl = 0; pl < kPlanes; ++pl, src += p.in_plane * 8)
                    bulk_g2s(dst + uint32_t(pl * p.in_pitch), src, uint32_t(p.run_bytes), bar);
            }
            __syncwarp();
        }
    } else if (warp <= 2) {
        // ============================================================ MMA issuers: warp 1 conv1, warp 2 conv2
        const uint32_t idesc = umma_idesc_f16(kTile, kC);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t in_lbo = (uint32_t(p.in_pitch) >> 4) << 16, in_kstep = (2u * uint32_t(p.in_pitch)) >> 4;
        constexpr uint32_t mid_lbo = (uint32_t(kMidPitch) >> 4) << 16, mid_kstep = (2u * uint32_t(kMidPitch)) >> 4;
        constexpr uint32_t b_lbo = (uint32_t(kSlab) >> 4) << 16, b_kstep = (2u * kSlab) >> 4, b_tap = (uint32_t(kPlanes) * kSlab) >> 4;
        uint32_t sh[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) sh[t] = uint32_t((t / 3 - 1) * Wp + (t % 3 - 1));
        const uint32_t w1_lo = (smem_u32(w_smem) >> 4) | b_lbo, w2_lo = ((smem_u32(w_smem) + kWBytes) >> 4) | b_lbo;
        MBAR_WAIT(smem_u32(&hdr->w_full), 0u);
        const bool prof = p.prof != nullptr;
        long long pt[5] = {0, 0, 0, 0, 0}, pt_mark = prof ? clock64() : 0;
        const long long pt_t0 = pt_mark;
        auto lap = [&](int k) { if (prof) { const long long t = clock64(); pt[k] += t - pt_mark; pt_mark = t; } };
        auto conv1 = [&](int j) {
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            MBAR_WAIT(smem_u32(&hdr->in_full[b]), par);
            lap(0);
            MBAR_WAIT(smem_u32(&hdr->acc1_free[b]), par ^ 1u);
            lap(1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(in_smem) + b * uint32_t(in_buf_bytes) + uint32_t(R) * 16u;
#pragma unroll
                for (int t = 0; t < kT1; ++t)
                    issue_linear<9, 2>(tmem_base + acc1_col(b, t), ((a0 + uint32_t(t) * (kTile * 16u)) >> 4) | in_lbo, w1_lo, sh,
                                       in_kstep, b_kstep, b_tap, desc_hi, idesc, 0u);
                umma_commit(smem_u32(&hdr->acc1_full[b]));
                umma_commit(smem_u32(&hdr->in_free[b]));
            }
            __syncwarp();
            lap(4);
        };
        auto conv2 = [&](int j) {
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            MBAR_WAIT(smem_u32(&hdr->mid_full[b]), par);
            lap(2);
            MBAR_WAIT(smem_u32(&hdr->acc2_free[b]), par ^ 1u);
            lap(3);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(mid_smem) + b * uint32_t(mid_buf_bytes) + uint32_t(kLead) * 16u;
#pragma unroll
                for (int g = 0; g < kG; ++g)
                    issue_linear<9, 2>(tmem_base + acc2_col(b, g), ((a0 + uint32_t(g) * (kTile * 16u)) >> 4) | mid_lbo, w2_lo, sh,
                                       mid_kstep, b_kstep, b_tap, desc_hi, idesc, 0u);
                umma_commit(smem_u32(&hdr->acc2_full[b]));
                umma_commit(smem_u32(&hdr->mid_free[b]));
            }
            __syncwarp();
            lap(4);
        };
        // (one warp issuing C1(0) C1(1) C2(0) C1(2) ... spent 47 cycles per MMA against the pipe's 40: two warps, each
        // with its own waits, keep the queue fed while the other one is between units)
        if (warp == 1) {
            for (int j = 0; j < my_units; ++j) conv1(j);
        } else {
            for (int j = 0; j < my_units; ++j) conv2(j);
        }
        if (prof && lane == 0) {
            unsigned long long* o = p.prof + (warp - 1) * 8;
            atomicAdd(o + 0, (unsigned long long)(clock64() - pt_t0));
            for (int k = 0; k < 5; ++k) atomicAdd(o + 1 + k, (unsigned long long)pt[k]);
            atomicAdd(o + 6, (unsigned long long)my_units);
            atomicAdd(o + 7, 1ull);
        }
    } else {
        // ============================================================ epilogue (8 warps)
        const int ew = warp - 3;
        const int set = ew >> 2;                        // tiles alternate between the two sets
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int row = lg * 32 + lane;
        const uint32_t lane_sel = uint32_t(lg * 32) << 16;
        // crop-relative position of padded-linear pixel q (any q > -HpWp): interior pixels are the real outputs
        auto interior_of = [&](long long q) {
            const uint32_t rem = uint32_t(q + HpWp) % uint32_t(HpWp);
            const uint32_t yy = rem / uint32_t(Wp), xx = rem - yy * uint32_t(Wp);
            return q >= 0 && q < p.P && yy >= 1u && yy <= uint32_t(p.H) && xx >= 1u && xx <= uint32_t(p.W);
        };
        auto epilogue1 = [&](int j) {
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long qm = unit * (kG * kTile) - kLead;               // pixel of row 0 of conv1 tile 0
            MBAR_WAIT(smem_u32(&hdr->acc1_full[b]), par);
            MBAR_WAIT(smem_u32(&hdr->mid_free[b]), par ^ 1u);               // C2 of the unit two back no longer reads the buffer
            tc_fence_after();
            uint8_t* mid = mid_smem + b * mid_buf_bytes;
            for (int t = set; t < kT1; t += 2) {
                uint32_t v[kC];
                const uint32_t taddr = tmem_base + acc1_col(b, t) + lane_sel;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                const bool keep = interior_of(qm + t * kTile + row);
                tmem_ld_wait();
                uint8_t* dst = mid + (t * kTile + row) * 16;
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl) {
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = keep ? fmaxf(__uint_as_float(v[pl * 8 + i]) + hdr->bias[0][pl * 8 + i], 0.f) : 0.f;
                    uint4 o4;
                    o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                    o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(dst + pl * kMidPitch) = o4;
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();               // these generic-proxy stores are read by tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&hdr->acc1_free[b]));
                mbar_arrive(smem_u32(&hdr->mid_full[b]));
            }
        };
        auto epilogue2 = [&](int j) {
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long q0 = unit * (kG * kTile) + row;
            const int g0 = (set + j) & 1;                                   // the set with two tiles alternates per unit
            // the residual = the block input at the output pixel: issued before the wait, it arrives under the MMAs
            uint4 res[2][kPlanes];
            bool keep[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int g = g0 + 2 * k;
                keep[k] = g < kG && interior_of(q0 + g * kTile);
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl)
                    res[k][pl] = keep[k] ? __ldg(reinterpret_cast<const uint4*>(p.in + ((long long)pl * p.in_plane + q0 + g * kTile) * 8))
                                         : make_uint4(0, 0, 0, 0);
            }
            MBAR_WAIT(smem_u32(&hdr->acc2_full[b]), par);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int g = g0 + 2 * k;
                if (g >= kG) break;
                uint32_t v[kC];
                const uint32_t taddr = tmem_base + acc2_col(b, g) + lane_sel;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                tmem_ld_wait();
                __half* outp = p.out + (q0 + g * kTile) * 8;
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl) {
                    const uint32_t rr[4] = {res[k][pl].x, res[k][pl].y, res[k][pl].z, res[k][pl].w};
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r2 = unpack_half2(rr[i]);
                        f[2 * i] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i]) + hdr->bias[1][pl * 8 + 2 * i] + r2.x, 0.f);
                        f[2 * i + 1] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i + 1]) + hdr->bias[1][pl * 8 + 2 * i + 1] + r2.y, 0.f);
                    }
                    if (keep[k]) {
                        uint4 o4;
                        o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                        o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                        *reinterpret_cast<uint4*>(outp + (long long)pl * p.out_plane * 8) = o4;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&hdr->acc2_free[b]));
        };
        for (int j = 0; j < my_units; ++j) {
            epilogue1(j);
            if (j > 0) epilogue2(j - 1);
        }
        if (my_units > 0) epilogue2(my_units - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

}  // namespace poco

using namespace poco;

extern "C" int poco_basic_block_supported(int32_t C, int32_t H, int32_t W) {
    // (a unit's input run starts kLead + R pixels before its first output pixel and ends kT1 * 128 - kLead + R pixels after
    // it, R = W + 3: both ends of the tensor must stay inside the activation guard)
    return C == kC && H >= 1 && W >= 1 && W + 3 <= kLead && (kT1 * kTile - kLead + W + 3) * 16 <= POCO_ACT_GUARD_BYTES;
}

extern "C" int poco_basic_block_run(const poco_basic_block* d, void* stream) {
    POCO_CHECK(d != nullptr, "null descriptor");
    if (check_act(d->in, "in") || check_act(d->out, "out")) return 1;
    const poco_act &in = d->in, &out = d->out;
    POCO_CHECK(in.C == out.C && in.N == out.N && in.H == out.H && in.W == out.W, "basic block: in and out must share one geometry");
    POCO_CHECK(poco_basic_block_supported(in.C, in.H, in.W), "basic block: only 32 channels with W + 3 <= 64 run fused");
    POCO_CHECK(in.lo == nullptr && out.lo == nullptr, "basic block: fp16 mode only");
    POCO_CHECK(in.data != out.data, "basic block: in and out must not alias");
    POCO_CHECK(d->weight1 && d->weight2 && d->bias1 && d->bias2, "null weight / bias");
    const int64_t P = int64_t(in.N) * (in.H + 2) * (in.W + 2);
    POCO_CHECK(P + 4096 < (int64_t(1) << 31), "tensor too large");
    const int R = in.W + 3;
    BlockParams p{};
    p.in = static_cast<const __half*>(in.data);
    p.out = static_cast<__half*>(out.data);
    p.w1 = static_cast<const __half*>(d->weight1);
    p.w2 = static_cast<const __half*>(d->weight2);
    p.b1 = d->bias1;
    p.b2 = d->bias2;
    p.in_plane = in.plane_stride;
    p.out_plane = out.plane_stride;
    p.H = in.H; p.W = in.W;
    p.P = int(P);
    p.num_units = int((P + kG * kTile - 1) / (kG * kTile));
    p.run_bytes = (kT1 * kTile + 2 * R) * 16;
    p.in_pitch = (p.run_bytes + 127) / 128 * 128;
    static const char* prof_env = getenv("POCO_BBLOCK_PROF");       // bring-up: device address (decimal) of 16 zeroed uint64 counters
    p.prof = prof_env ? reinterpret_cast<unsigned long long*>(strtoull(prof_env, nullptr, 10)) : nullptr;
    const size_t smem = size_t(kHeader) + 2 * kWBytes + 2 * size_t(kPlanes) * p.in_pitch + 2 * size_t(kPlanes) * kMidPitch;
    POCO_CHECK(smem <= 227 * 1024, "shared memory");
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(basic_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const int budget = d->max_ctas > 0 ? std::min(d->max_ctas, sm_count()) : sm_count();
    const int grid = std::max(1, std::min(p.num_units, budget));
    basic_block_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    POCO_LAUNCHED();
    return 0;
}
