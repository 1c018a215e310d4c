// poco_b200 -- the tail of a Bottleneck block as ONE tcgen05 launch (sm_100a):
//     out = ReLU(BN3(conv3_1x1(ReLU(BN2(conv2_3x3(in))))) + residual)        (hrnet.py:79-99, resnet.py:100-121)
// for layer1 of the HRNet trunks (hrnet.py:306, planes = 64: 64 -> 64 -> 256 channels at 56x56).  As two poco_conv
// launches the 64-channel intermediate makes an HBM round trip and the 1x1 conv -- 4 K steps per tile -- is all
// pipeline fill and epilogue (204 us per block at batch 256 against 142 us of HBM time).  Here a work unit is ONE
// 128-pixel tile: the 3x3 conv's accumulator goes through epilogue 1 (shift2, ReLU, fp16) into shared memory in the
// operand layout, the 1x1 conv -- pointwise, so no halo and no recompute -- reads it there as two N = 128 halves, and
// epilogue 2 adds shift3 and the 256-channel residual and writes the output planes.
//
// Roles (96 + 128 kSets threads, one persistent CTA per SM, tiles dealt round-robin): warp 0 producer (both weight tensors once,
// then one bulk copy per input plane and tile: the tile plus its (W+3)-pixel reach, double buffered), warp 1 issues the
// 3x3 conv's MMAs (9 taps x 4 K steps, N = 64), warp 2 the 1x1 conv's (4 K steps per N = 128 half), then 4 * kSets
// epilogue warps: kSets sets of four TMEM lane groups; set s owns columns [64 s / kSets, ...) of the first accumulator and
// output channels [256 s / kSets, ...) of the second.  TMEM: 2 x 64 + 2 x 128 columns.  The kernel is bound by its 822 MB of
// output + residual traffic -- and by the latency of its epilogues: with FOUR sets (16 epilogue warps, 88 registers, one
// TMEM buffer per thread) instead of two (150 registers, double buffered) a launch takes 203 instead of 220 us at batch 256
// (4.55 TB/s; the whole step 10.70 -> 10.58 ms, A/B twice on one box).
//
// Measured while tuning (batch 256, 45.5 tiles per CTA, clock64 laps inside the epilogue warps, tools/btail_bench.py of the
// time): a tile takes ~8200 cycles whatever the role layout -- 218 us per launch = 4.2 TB/s of algorithmic traffic, DRAM
// 51 % busy (ncu).  Not kept: requesting tile j + 1's residual rows while tile j is written (231 us); ONE issuer warp in
// the order S2(j), S1(j + 2) (238 us); four issuer warps taking turns so that no single SM sub-partition hosts every
// tcgen05.mma (231 us).  That last experiment did show that the epilogue warps which share a sub-partition with the
// issuing warp are the slow ones (5700 against 3800 cycles of epilogue 2 per tile, and the slow lane group moves with
// the issuer warp): their tcgen05.ld queue behind the MMA batches issued next to them.
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

constexpr int kCm = 64;                 // channels of the input and of the intermediate
constexpr int kCo = 256;                // output channels
constexpr int kPm = kCm / 8, kPo = kCo / 8;
constexpr int kTile = 128;
constexpr int kSlab2 = kCm * 16;        // 3x3 conv: one (tap, 8-channel) weight slab = 64 output channels x 16 B
constexpr int kSlab3 = kCo * 16;        // 1x1 conv: 256 output channels x 16 B
constexpr int kW2Bytes = 9 * kPm * kSlab2;          // 72 KB
constexpr int kW3Bytes = kPm * kSlab3;              // 32 KB
constexpr int kMidPitch = kTile * 16;               // one plane of the intermediate tile: 2 KB
constexpr int kHeader = 2048;
// epilogue warp sets (four TMEM lane groups each).  The epilogues are latency-bound, not TMEM-bound (tools/tmem_ld_bench.cu:
// a 32-column tcgen05.ld takes ~190 cycles, eight warps with several loads in flight read > 600 B/clk): with two sets a
// scheduler holds two epilogue warps and cannot hide their tcgen05.ld / LDS / LDG latencies, four sets give it four.
#ifndef POCO_TAIL_SETS
#define POCO_TAIL_SETS 4
#endif
constexpr int kSets = POCO_TAIL_SETS;                   // 2 or 4
static_assert(kSets == 2 || kSets == 4, "POCO_TAIL_SETS");
constexpr int kThreads = 96 + 128 * kSets;              // producer, two issuers, 4 * kSets epilogue warps
constexpr int kCols1 = kCm / kSets;                     // epilogue 1: accumulator columns per set (32 or 16)
constexpr int kCols2 = kCo / kSets;                     // epilogue 2: output channels per set (128 or 64)
constexpr int kTmemCols = 512;          // 2 x 64 + 2 x 128 = 384 used

struct TailParams {
    const __half* in;
    __half* out;
    const __half* res;
    const __half* w2;
    const __half* w3;
    const float* b2;
    const float* b3;
    long long in_plane, out_plane, res_plane;       // plane strides in pixels
    int H, W;
    int P;                              // N * (H + 2) * (W + 2)
    int num_units;                      // tiles
    int in_pitch;                       // bytes between the planes of an input run in shared memory
    int run_bytes;                      // bytes of one input run: (128 + 2 R) * 16
};

struct Header {
    unsigned long long in_full[2], in_free[2], acc1_full[2], acc1_free[2], mid_full[2], mid_free[2], acc2_full[2], acc2_free[2];
    unsigned long long w_full;
    uint32_t tmem_base;
    uint32_t pad_;
    float bias2[kCm];
    float bias3[kCo];
};
static_assert(sizeof(Header) <= kHeader, "header too large");

__global__ void __launch_bounds__(kThreads, 1) bottleneck_tail_kernel(const TailParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Header* hdr = reinterpret_cast<Header*>(smem);
    uint8_t* w2_smem = smem + kHeader;                              // [tap][plane][64][8]
    uint8_t* w3_smem = w2_smem + kW2Bytes;                          // [plane][256][8]
    uint8_t* in_smem = w3_smem + kW3Bytes;                          // [2][plane][run]
    const int in_buf_bytes = kPm * p.in_pitch;
    uint8_t* mid_smem = in_smem + 2 * in_buf_bytes;                 // [2][plane][128 pixels]
    constexpr int mid_buf_bytes = kPm * kMidPitch;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int Wp = p.W + 2, HpWp = (p.H + 2) * Wp, R = Wp + 1;
    const int my_units = (p.num_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

    if (threadIdx.x < 16) {
        unsigned long long* bars = hdr->in_full;                    // the sixteen ring barriers are contiguous
        const int which = threadIdx.x >> 1;                         // 0 in_full 1 in_free 2 acc1_full 3 acc1_free 4 mid_full 5 mid_free 6 acc2_full 7 acc2_free
        const uint32_t count = (which == 3 || which == 4) ? uint32_t(4 * kSets) : (which == 7 ? uint32_t(2 * kSets) : 1u);    // all epilogue warps / the warps of an accumulator half / one commit
        mbar_init(smem_u32(bars + threadIdx.x), count);
    }
    if (threadIdx.x == 16) mbar_init(smem_u32(&hdr->w_full), 1);
    if (threadIdx.x >= 32 && threadIdx.x < 32 + kCm) hdr->bias2[threadIdx.x - 32] = p.b2[threadIdx.x - 32];
    if (threadIdx.x >= 96 && threadIdx.x < 96 + kCo) hdr->bias3[threadIdx.x - 96] = p.b3[threadIdx.x - 96];
    mbar_fence_init();
    if (warp == 1) tmem_alloc(smem_u32(&hdr->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    auto acc1_col = [&](uint32_t b) { return b * uint32_t(kCm); };
    auto acc2_col = [&](uint32_t h) { return 2u * kCm + h * 128u; };

    if (warp == 0) {
        // ============================================================ producer
        if (elect_one()) {
            const uint32_t bar = smem_u32(&hdr->w_full);
            mbar_arrive_expect_tx(bar, uint32_t(kW2Bytes + kW3Bytes));
            bulk_g2s(smem_u32(w2_smem), p.w2, kW2Bytes, bar);
            bulk_g2s(smem_u32(w3_smem), p.w3, kW3Bytes, bar);
        }
        __syncwarp();
        for (int j = 0; j < my_units; ++j) {
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            MBAR_WAIT(smem_u32(&hdr->in_free[b]), par ^ 1u);        // the 3x3 MMAs of the tile two back have retired
            if (elect_one()) {
                const uint32_t bar = smem_u32(&hdr->in_full[b]);
                mbar_arrive_expect_tx(bar, uint32_t(kPm) * uint32_t(p.run_bytes));
                const long long q0 = unit * kTile - R;
                const __half* src = p.in + q0 * 8;
                const uint32_t dst = smem_u32(in_smem) + b * uint32_t(in_buf_bytes);
                for (int pl = 0; pl < kPm; ++pl, src += p.in_plane * 8)
                    bulk_g2s(dst + uint32_t(pl * p.in_pitch), src, uint32_t(p.run_bytes), bar);
            }
            __syncwarp();
        }
    } else if (warp <= 2) {
        // ============================================================ MMA issuers: warp 1 the 3x3 conv, warp 2 the 1x1 conv
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        MBAR_WAIT(smem_u32(&hdr->w_full), 0u);
        if (warp == 1) {
            const uint32_t idesc = umma_idesc_f16(kTile, kCm);
            const uint32_t in_lbo = (uint32_t(p.in_pitch) >> 4) << 16, in_kstep = (2u * uint32_t(p.in_pitch)) >> 4;
            constexpr uint32_t b_lbo = (uint32_t(kSlab2) >> 4) << 16, b_kstep = (2u * kSlab2) >> 4, b_tap = (uint32_t(kPm) * kSlab2) >> 4;
            uint32_t sh[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) sh[t] = uint32_t((t / 3 - 1) * Wp + (t % 3 - 1));
            const uint32_t w_lo = (smem_u32(w2_smem) >> 4) | b_lbo;
            for (int j = 0; j < my_units; ++j) {
                const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
                MBAR_WAIT(smem_u32(&hdr->in_full[b]), par);
                MBAR_WAIT(smem_u32(&hdr->acc1_free[b]), par ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(in_smem) + b * uint32_t(in_buf_bytes) + uint32_t(R) * 16u;
                    issue_linear<9, 4>(tmem_base + acc1_col(b), (a0 >> 4) | in_lbo, w_lo, sh, in_kstep, b_kstep, b_tap, desc_hi, idesc, 0u);
                    umma_commit(smem_u32(&hdr->acc1_full[b]));
                    umma_commit(smem_u32(&hdr->in_free[b]));
                }
                __syncwarp();
            }
        } else {
            const uint32_t idesc = umma_idesc_f16(kTile, 128);
            constexpr uint32_t mid_lbo = (uint32_t(kMidPitch) >> 4) << 16, mid_kstep = (2u * uint32_t(kMidPitch)) >> 4;
            constexpr uint32_t b_lbo = (uint32_t(kSlab3) >> 4) << 16, b_kstep = (2u * kSlab3) >> 4;
            const uint32_t sh[9] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            for (int j = 0; j < my_units; ++j) {
                const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u, upar = uint32_t(j) & 1u;
                MBAR_WAIT(smem_u32(&hdr->mid_full[b]), par);
                const uint32_t a_lo = ((smem_u32(mid_smem) + b * uint32_t(mid_buf_bytes)) >> 4) | mid_lbo;
#pragma unroll
                for (int h = 0; h < 2; ++h) {           // half h = output channels [128 h, 128 h + 128): its own accumulator and barriers
                    MBAR_WAIT(smem_u32(&hdr->acc2_free[h]), upar ^ 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w_lo = ((smem_u32(w3_smem) + uint32_t(h) * 128u * 16u) >> 4) | b_lbo;
                        issue_linear<1, 4>(tmem_base + acc2_col(uint32_t(h)), a_lo, w_lo, sh, mid_kstep, b_kstep, 0u, desc_hi, idesc, 0u);
                        umma_commit(smem_u32(&hdr->acc2_full[h]));
                        if (h == 1) umma_commit(smem_u32(&hdr->mid_free[b]));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ============================================================ epilogue (4 * kSets warps)
        const int ew = warp - 3;
        const int set = ew >> 2;
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int row = lg * 32 + lane;
        const uint32_t lane_sel = uint32_t(lg * 32) << 16;
        auto interior_of = [&](long long q) {
            const uint32_t rem = uint32_t(q) % uint32_t(HpWp);
            const uint32_t yy = rem / uint32_t(Wp), xx = rem - yy * uint32_t(Wp);
            return q < p.P && yy >= 1u && yy <= uint32_t(p.H) && xx >= 1u && xx <= uint32_t(p.W);
        };
        auto epilogue1 = [&](int j) {                   // this warp: columns [kCols1 set, kCols1 set + kCols1) of its 32 rows
            const uint32_t b = uint32_t(j) & 1u, par = (uint32_t(j) >> 1) & 1u;
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            MBAR_WAIT(smem_u32(&hdr->acc1_full[b]), par);
            MBAR_WAIT(smem_u32(&hdr->mid_free[b]), par ^ 1u);
            tc_fence_after();
            uint32_t v[kCols1];
            const uint32_t taddr = tmem_base + acc1_col(b) + uint32_t(set * kCols1) + lane_sel;
#pragma unroll
            for (int c = 0; c < kCols1; c += 16) tmem_ld16(taddr + uint32_t(c), v + c);
            const bool keep = interior_of(unit * kTile + row);
            tmem_ld_wait();
            uint8_t* dst = mid_smem + b * mid_buf_bytes + (set * (kCols1 / 8)) * kMidPitch + row * 16;
#pragma unroll
            for (int pl = 0; pl < kCols1 / 8; ++pl) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = keep ? fmaxf(__uint_as_float(v[pl * 8 + i]) + hdr->bias2[set * kCols1 + pl * 8 + i], 0.f) : 0.f;
                uint4 o4;
                o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                *reinterpret_cast<uint4*>(dst + pl * kMidPitch) = o4;
            }
            tc_fence_before();
            fence_proxy_async_smem();               // these generic-proxy stores are read by tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&hdr->acc1_free[b]));
                mbar_arrive(smem_u32(&hdr->mid_full[b]));
            }
        };
        constexpr int kPl2 = kCols2 / 8;                // output planes per set (16 or 8)
        auto res_ptr = [&](long long q, int pl) {
            return reinterpret_cast<const uint4*>(p.res + ((long long)(set * kPl2 + pl) * p.res_plane + q) * 8);
        };
        const long long q_first = (long long)blockIdx.x * kTile + row;
        const int half = (set * kCols2) / 128;          // which N = 128 accumulator half this set reads
        auto epilogue2 = [&](int j) {                   // this warp: output channels [kCols2 set, kCols2 set + kCols2) of its 32 rows
            const uint32_t upar = uint32_t(j) & 1u;
            const long long q = q_first + (long long)j * gridDim.x * kTile;
            const bool keep = interior_of(q);
            // the residual rows of all planes are requested before the wait: they arrive under the MMAs
            uint4 res[kPl2];
#pragma unroll
            for (int pl = 0; pl < kPl2; ++pl) res[pl] = keep ? __ldg(res_ptr(q, pl)) : make_uint4(0, 0, 0, 0);
            MBAR_WAIT(smem_u32(&hdr->acc2_full[half]), upar);
            tc_fence_after();
            __half* outp = p.out + ((long long)(set * kPl2) * p.out_plane + q) * 8;
            // accumulator columns 32 at a time (4 output planes); the TMEM load of the next 32 is in flight during the math
            // (two sets: the next load is in flight during the math of this one; four sets leave a thread 104 registers, so one
            //  buffer, and the other three warps of the scheduler cover the load)
            constexpr bool kDouble = kSets == 2;
            uint32_t va[32], vb[kDouble ? 32 : 1];
            const uint32_t taddr0 = tmem_base + acc2_col(uint32_t(half)) + uint32_t((set * kCols2) % 128) + lane_sel;
            tmem_ld16(taddr0, va);
            tmem_ld16(taddr0 + 16, va + 16);
            constexpr int kChunks = kCols2 / 32;
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                uint32_t* v = (kDouble && (c & 1)) ? vb : va;
                uint32_t* vn = (kDouble && !(c & 1)) ? vb : va;
                if (!kDouble && c > 0) {
                    tmem_ld16(taddr0 + uint32_t(c * 32), va);
                    tmem_ld16(taddr0 + uint32_t(c * 32 + 16), va + 16);
                }
                tmem_ld_wait();
                if (kDouble && c < kChunks - 1) {
                    tmem_ld16(taddr0 + uint32_t((c + 1) * 32), vn);
                    tmem_ld16(taddr0 + uint32_t((c + 1) * 32 + 16), vn + 16);
                }
                if (c == kChunks - 1) {     // this set's columns are drained: once every set of the half has arrived, the next tile's MMAs may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&hdr->acc2_free[half]));
                }
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const uint4 r4 = res[c * 4 + pl];
                    const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
                    const float* bs = hdr->bias3 + set * kCols2 + c * 32 + pl * 8;
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r2 = unpack_half2(rr[i]);
                        f[2 * i] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i]) + bs[2 * i] + r2.x, 0.f);
                        f[2 * i + 1] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i + 1]) + bs[2 * i + 1] + r2.y, 0.f);
                    }
                    if (keep) {
                        uint4 o4;
                        o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                        o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                        *reinterpret_cast<uint4*>(outp + (long long)(c * 4 + pl) * p.out_plane * 8) = o4;
                    }
                }
            }
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

extern "C" int poco_bottleneck_tail_supported(int32_t Cmid, int32_t Cout, int32_t H, int32_t W) {
    const int in_pitch = ((kTile + 2 * (W + 3)) * 16 + 127) / 128 * 128;
    return Cmid == kCm && Cout == kCo && H >= 1 && W >= 1 && (kTile + W + 3) * 16 <= POCO_ACT_GUARD_BYTES &&
           kHeader + kW2Bytes + kW3Bytes + 2 * kPm * in_pitch + 2 * kPm * kMidPitch <= 227 * 1024;
}

extern "C" int poco_bottleneck_tail_run(const poco_bottleneck_tail* d, void* stream) {
    POCO_CHECK(d != nullptr, "null descriptor");
    if (check_act(d->in, "in") || check_act(d->out, "out")) return 1;
    const poco_act &in = d->in, &out = d->out;
    POCO_CHECK(in.N == out.N && in.H == out.H && in.W == out.W, "bottleneck tail: in and out must share one geometry");
    POCO_CHECK(poco_bottleneck_tail_supported(in.C, out.C, in.H, in.W), "bottleneck tail: only 64 -> 64 -> 256 channels run fused");
    POCO_CHECK(in.lo == nullptr && out.lo == nullptr, "bottleneck tail: fp16 mode only");
    POCO_CHECK(d->weight2 && d->weight3 && d->bias2 && d->bias3 && d->residual, "null weight / bias / residual");
    const int64_t P = int64_t(in.N) * (in.H + 2) * (in.W + 2);
    POCO_CHECK(d->res_plane_stride >= P, "residual plane stride too small");
    POCO_CHECK(d->residual != out.data, "bottleneck tail: the residual must not alias the output");
    POCO_CHECK(P + 4096 < (int64_t(1) << 31), "tensor too large");
    const int R = in.W + 3;
    TailParams p{};
    p.in = static_cast<const __half*>(in.data);
    p.out = static_cast<__half*>(out.data);
    p.res = static_cast<const __half*>(d->residual);
    p.w2 = static_cast<const __half*>(d->weight2);
    p.w3 = static_cast<const __half*>(d->weight3);
    p.b2 = d->bias2;
    p.b3 = d->bias3;
    p.in_plane = in.plane_stride;
    p.out_plane = out.plane_stride;
    p.res_plane = d->res_plane_stride;
    p.H = in.H; p.W = in.W;
    p.P = int(P);
    p.num_units = int((P + kTile - 1) / kTile);
    p.run_bytes = (kTile + 2 * R) * 16;
    p.in_pitch = (p.run_bytes + 127) / 128 * 128;
    const size_t smem = size_t(kHeader) + kW2Bytes + kW3Bytes + 2 * size_t(kPm) * p.in_pitch + 2 * size_t(kPm) * kMidPitch;
    POCO_CHECK(smem <= 227 * 1024, "shared memory");
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(bottleneck_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const int budget = d->max_ctas > 0 ? std::min(d->max_ctas, sm_count()) : sm_count();
    const int grid = std::max(1, std::min(p.num_units, budget));
    bottleneck_tail_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    POCO_LAUNCHED();
    return 0;
}
