// poco_b200 -- the 64-channel residual BasicBlock as ONE tcgen05 launch with conv2's weights STREAMED (sm_100a):
//     out = ReLU(BN2(conv2(ReLU(BN1(conv1(in))))) + in)          (hrnet.py:42-58, downsample is None)
// for the 64-channel 28x28 branch of the HRNet modules (hrnet.py:140-186): 64 of the 311 conv launches of a POCO-CLIFF /
// HRNet-W32 forward and, with the lanes of a module running side by side, the branch that ends most modules.
//
// bblock_tc.cu keeps both weight tensors of a block in shared memory; at 64 channels they take 144 KB, which leaves room
// for ONE conv2 tile per work unit (two conv1 tiles: 1.5x the MMAs of the two-launch block) and that flavour measured
// slower than the two launches.  Here only conv1's weights are resident (72 KB); conv2's stream through a four- to eight-deep
// ring of 8 KB filter-tap stages (72 KB per unit out of L2, fetched by a warp of its own).  That buys a unit of G = 2 conv2
// tiles (256 output pixels) on three conv1 tiles: 1.25x the MMAs, one input run of 384 + 2 (W + 3) pixels per plane.
//
//   work unit u: conv2 tiles [u * 256, u * 256 + 256) of the padded-linear pixel axis; conv1 on [u * 256 - 64, + 384).
//   C1(u): 3 tiles x 36 MMAs (M = 128, N = 64, K = 16: 48 cycles each) from the landed input run, one accumulator per
//          tile, committed tile by tile so that epilogue 1 starts under the MMAs of the next tile;
//   E1(u): fp16 ReLU(acc + shift1), zero at halo / out-of-range pixels, into the intermediate buffer (operand layout);
//   C2(u): filter tap by filter tap as the stages land: both conv2 tiles consume a stage (8 MMAs), then it is released;
//   E2(u): + shift2 + residual (the block input, re-read through L2), ReLU, 16-byte pixels of the eight output planes.
// Shared memory (single buffers): W1 72 KB | ring 32 KB | input run <= 64 KB | intermediate 48 KB.  TMEM: conv1's three
// accumulators double buffered (384 columns), conv2's two single buffered (128).  On the tensor pipe the order is
// C1(u) C2(u) C1(u + 1) ... (conv1's issuer waits until conv2's has queued the previous unit): the next input run lands
// chunk by chunk as C1(u)'s tiles retire, E1(u) converts tile by tile under C1(u)'s later tiles, E2(u) runs under
// C1(u + 1), and the pipe idles only while the last tile of E1(u) is converted.
//
// Roles (384 threads, one persistent CTA per SM, units dealt round-robin): warp 0 producer (W1 once, then the input
// runs), warp 1 issues conv1, warp 2 conv2, warp 3 streams W2, warps 4-11 epilogue (two sets of four TMEM lane groups;
// a set takes one 32-column half of every tile).
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

constexpr int kC = 64;
constexpr int kPlanes = kC / 8;
constexpr int kTile = 128;
constexpr int kG = 2;                       // conv2 tiles per unit
constexpr int kT1 = kG + 1;                 // conv1 tiles per unit
constexpr int kLead = 64;                   // conv1 starts this many pixels before the unit's first output pixel (>= W + 3)
constexpr int kHeader = 1024;
constexpr int kThreads = 384;
constexpr int kSlab = kC * 16;                          // one (tap, 8-channel) weight slab: 64 output channels x 16 B
constexpr int kTapBytes = kPlanes * kSlab;              // one filter tap of a conv: 8 KB = one ring stage
constexpr int kWBytes = 9 * kTapBytes;                  // one conv's weights
constexpr int kRing = 4;                    // minimum depth of the W2 ring (the launcher spends spare shared memory on more stages)
constexpr int kMaxRing = 8;
constexpr int kMidPitch = kT1 * kTile * 16;             // one plane of the intermediate
constexpr int kMidBytes = kPlanes * kMidPitch;
constexpr int kTmemCols = 512;
static_assert(2 * kT1 * kC + kG * kC <= kTmemCols, "accumulators do not fit in TMEM");

struct Block64Params {
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
    int run_bytes;                      // bytes of one input run: (3 * 128 + 2 R) * 16
    int ring;                           // W2 ring stages (kRing..kMaxRing)
    unsigned long long* prof;           // bring-up (POCO_BBLOCK_PROF): per issuer [total, wait a, wait b, wait c, issue, -, units, ctas]
};

struct Header {
    unsigned long long w1_full, in_full[kT1], in_free[kT1], mid_full, mid_free, acc2_full, acc2_free, c2_issued;
    unsigned long long acc1_full[2][kT1], acc1_free[2];
    unsigned long long w2_full[kMaxRing], w2_empty[kMaxRing];
    uint32_t tmem_base;
    uint32_t pad_;
    float bias[2][kC];
};
static_assert(sizeof(Header) <= kHeader, "header too large");

__global__ void __launch_bounds__(kThreads, 1) basic_block64_kernel(const Block64Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Header* hdr = reinterpret_cast<Header*>(smem);
    uint8_t* w1_smem = smem + kHeader;                              // [tap][plane][64][8]
    uint8_t* ring = w1_smem + kWBytes;                              // [stage][plane][64][8]
    uint8_t* in_smem = ring + p.ring * kTapBytes;                   // [plane][run]
    uint8_t* mid_smem = in_smem + kPlanes * p.in_pitch;             // [plane][3 * 128 pixels]

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int Wp = p.W + 2, HpWp = (p.H + 2) * Wp, R = Wp + 1;
    const int my_units = (p.num_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&hdr->w1_full), 1);
        for (int c = 0; c < kT1; ++c) {
            mbar_init(smem_u32(&hdr->in_full[c]), 1);
            mbar_init(smem_u32(&hdr->in_free[c]), 1);       // conv1's commit behind tile c
        }
        mbar_init(smem_u32(&hdr->mid_full), 8);             // the eight epilogue warps
        mbar_init(smem_u32(&hdr->mid_free), 1);             // conv2's commit
        mbar_init(smem_u32(&hdr->acc2_full), 1);
        mbar_init(smem_u32(&hdr->acc2_free), 8);
        mbar_init(smem_u32(&hdr->c2_issued), 1);            // conv2's issuer, once its last MMA of a unit is queued
        for (int b = 0; b < 2; ++b) {
            for (int t = 0; t < kT1; ++t) mbar_init(smem_u32(&hdr->acc1_full[b][t]), 1);
            mbar_init(smem_u32(&hdr->acc1_free[b]), 8);
        }
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(smem_u32(&hdr->w2_full[s]), 1);
            mbar_init(smem_u32(&hdr->w2_empty[s]), 1);
        }
    }
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
    auto acc2_col = [&](int g) { return (2u * kT1 + uint32_t(g)) * uint32_t(kC); };

    if (warp == 0) {
        // ============================================================ producer: W1 once, then the input runs
        if (elect_one()) {
            const uint32_t bar = smem_u32(&hdr->w1_full);
            mbar_arrive_expect_tx(bar, uint32_t(kWBytes));
            bulk_g2s(smem_u32(w1_smem), p.w1, kWBytes, bar);
        }
        __syncwarp();
        // The run of a unit lands in three chunks per plane -- pixels [0, 128), [128, 256), [256, 384 + 2 R) -- and conv1 tile t
        // reads run pixels [128 t, 128 t + 128 + 2 R): chunk c is free again as soon as tile c has retired, so the first two
        // chunks of the NEXT unit's run are fetched under the remaining tiles of this one instead of after all of them.
        for (int j = 0; j < my_units; ++j) {
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long q0 = unit * (kG * kTile) - kLead - R;           // (>= -8 KB guard, see the launcher)
            for (int c = 0; c < kT1; ++c) {
                MBAR_WAIT(smem_u32(&hdr->in_free[c]), uint32_t(j & 1) ^ 1u);        // tile c of the previous unit has retired
                if (elect_one()) {
                    const uint32_t bar = smem_u32(&hdr->in_full[c]);
                    const uint32_t bytes = c < kT1 - 1 ? uint32_t(kTile * 16) : uint32_t(p.run_bytes - (kT1 - 1) * kTile * 16);
                    mbar_arrive_expect_tx(bar, uint32_t(kPlanes) * bytes);
                    const __half* src = p.in + (q0 + c * kTile) * 8;
                    for (int pl = 0; pl < kPlanes; ++pl, src += p.in_plane * 8)
                        bulk_g2s(smem_u32(in_smem) + uint32_t(pl * p.in_pitch + c * kTile * 16), src, bytes, bar);
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ============================================================ W2 stream: nine tap stages per unit
        uint32_t slot = 0, par = 0;
        const uint8_t* w2 = reinterpret_cast<const uint8_t*>(p.w2);
        for (int j = 0; j < my_units; ++j) {
            for (int tap = 0; tap < 9; ++tap) {
                MBAR_WAIT(smem_u32(&hdr->w2_empty[slot]), par ^ 1u);
                if (elect_one()) {
                    const uint32_t bar = smem_u32(&hdr->w2_full[slot]);
                    mbar_arrive_expect_tx(bar, uint32_t(kTapBytes));
                    bulk_g2s(smem_u32(ring) + slot * uint32_t(kTapBytes), w2 + size_t(tap) * kTapBytes, uint32_t(kTapBytes), bar);
                }
                __syncwarp();
                if (++slot == uint32_t(p.ring)) { slot = 0; par ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ============================================================ conv1 issuer
        const uint32_t idesc = umma_idesc_f16(kTile, kC);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t in_lbo = (uint32_t(p.in_pitch) >> 4) << 16, in_kstep = (2u * uint32_t(p.in_pitch)) >> 4;
        constexpr uint32_t b_lbo = (uint32_t(kSlab) >> 4) << 16, b_kstep = (2u * kSlab) >> 4, b_tap = uint32_t(kTapBytes) >> 4;
        uint32_t sh[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) sh[t] = uint32_t((t / 3 - 1) * Wp + (t % 3 - 1));
        const uint32_t w1_lo = (smem_u32(w1_smem) >> 4) | b_lbo;
        MBAR_WAIT(smem_u32(&hdr->w1_full), 0u);
        const bool prof = p.prof != nullptr;
        long long pt[4] = {0, 0, 0, 0}, pt_mark = prof ? clock64() : 0;
        const long long pt_t0 = pt_mark;
        auto lap = [&](int k) { if (prof) { const long long t = clock64(); pt[k] += t - pt_mark; pt_mark = t; } };
        for (int j = 0; j < my_units; ++j) {
            const uint32_t ab = uint32_t(j) & 1u, ap = (uint32_t(j) >> 1) & 1u;
            MBAR_WAIT(smem_u32(&hdr->in_full[0]), uint32_t(j & 1));
            MBAR_WAIT(smem_u32(&hdr->in_full[1]), uint32_t(j & 1));
            lap(0);
            MBAR_WAIT(smem_u32(&hdr->acc1_free[ab]), ap ^ 1u);
            // The tensor pipe runs MMAs in issue order.  C1(j) goes in BEHIND C2(j - 1): issued side by side the two finish
            // together, and E1(j) -- which overwrites the single intermediate buffer C2(j - 1) reads -- could not start under
            // C1(j)'s own tiles; the pipe then idled through E1 and E2 every unit (measured 11.7 k cycles per unit for 8.6 k of MMAs).
            if (j > 0) MBAR_WAIT(smem_u32(&hdr->c2_issued), uint32_t(j - 1) & 1u);
            lap(1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(in_smem) + uint32_t(R) * 16u;
#pragma unroll
            for (int t = 0; t < kT1; ++t) {
                if (t == 1) {                       // tile 1 reaches into the last chunk (and tile 2 ends in it)
                    MBAR_WAIT(smem_u32(&hdr->in_full[2]), uint32_t(j & 1));
                    lap(0);
                }
                if (elect_one()) {
                    issue_linear<9, kC / 16>(tmem_base + acc1_col(ab, t), ((a0 + uint32_t(t) * (kTile * 16u)) >> 4) | in_lbo, w1_lo, sh,
                                             in_kstep, b_kstep, b_tap, desc_hi, idesc, 0u);
                    umma_commit(smem_u32(&hdr->acc1_full[ab][t]));
                    umma_commit(smem_u32(&hdr->in_free[t]));
                }
                __syncwarp();
            }
            __syncwarp();
            lap(3);
        }
        if (prof && lane == 0) {
            atomicAdd(p.prof + 0, (unsigned long long)(clock64() - pt_t0));
            for (int k = 0; k < 4; ++k) atomicAdd(p.prof + 1 + k, (unsigned long long)pt[k]);
            atomicAdd(p.prof + 6, (unsigned long long)my_units);
            atomicAdd(p.prof + 7, 1ull);
        }
    } else if (warp == 2) {
        // ============================================================ conv2 issuer: tap by tap as the stages land
        const uint32_t idesc = umma_idesc_f16(kTile, kC);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        constexpr uint32_t mid_lbo = (uint32_t(kMidPitch) >> 4) << 16, mid_kstep = (2u * uint32_t(kMidPitch)) >> 4;
        constexpr uint32_t b_lbo = (uint32_t(kSlab) >> 4) << 16, b_kstep = (2u * kSlab) >> 4;
        const uint32_t mid16 = (smem_u32(mid_smem) >> 4) + uint32_t(kLead);
        const uint32_t ring16 = smem_u32(ring) >> 4;
        uint32_t slot = 0, par = 0;
        const bool prof = p.prof != nullptr;
        long long pt[4] = {0, 0, 0, 0}, pt_mark = prof ? clock64() : 0;
        const long long pt_t0 = pt_mark;
        auto lap = [&](int k) { if (prof) { const long long t = clock64(); pt[k] += t - pt_mark; pt_mark = t; } };
        for (int j = 0; j < my_units; ++j) {
            MBAR_WAIT(smem_u32(&hdr->mid_full), uint32_t(j & 1));
            lap(0);
            MBAR_WAIT(smem_u32(&hdr->acc2_free), uint32_t(j & 1) ^ 1u);
            lap(1);
            tc_fence_after();
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const uint32_t shift = uint32_t((tap / 3 - 1) * Wp + (tap % 3 - 1));
                MBAR_WAIT(smem_u32(&hdr->w2_full[slot]), par);
                lap(2);
                if (elect_one()) {
                    const uint32_t b0 = (ring16 + slot * (uint32_t(kTapBytes) >> 4)) | b_lbo;
#pragma unroll
                    for (int g = 0; g < kG; ++g) {
                        const uint32_t a0 = ((mid16 + uint32_t(g * kTile) + shift) & 0x3FFFu) | mid_lbo;
#pragma unroll
                        for (int k = 0; k < kC / 16; ++k)
                            umma_f16(tmem_base + acc2_col(g), desc64(desc_hi, a0 + uint32_t(k) * mid_kstep), desc64(desc_hi, b0 + uint32_t(k) * b_kstep),
                                     idesc, (tap | k) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&hdr->w2_empty[slot]));
                    if (tap == 8) {
                        umma_commit(smem_u32(&hdr->acc2_full));
                        umma_commit(smem_u32(&hdr->mid_free));
                        mbar_arrive(smem_u32(&hdr->c2_issued));
                    }
                }
                __syncwarp();
                lap(3);
                if (++slot == uint32_t(p.ring)) { slot = 0; par ^= 1u; }
            }
        }
        if (prof && lane == 0) {
            atomicAdd(p.prof + 8, (unsigned long long)(clock64() - pt_t0));
            for (int k = 0; k < 4; ++k) atomicAdd(p.prof + 9 + k, (unsigned long long)pt[k]);
            atomicAdd(p.prof + 14, (unsigned long long)my_units);
            atomicAdd(p.prof + 15, 1ull);
        }
    } else {
        // ============================================================ epilogue (8 warps)
        const int set = (warp - 4) >> 2;                // which 32-column half of a tile this warp converts
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
            const uint32_t ab = uint32_t(j) & 1u, ap = (uint32_t(j) >> 1) & 1u;
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long qm = unit * (kG * kTile) - kLead;               // pixel of row 0 of conv1 tile 0
            MBAR_WAIT(smem_u32(&hdr->mid_free), uint32_t(j & 1) ^ 1u);      // C2 of the previous unit has retired
            const float* bs = hdr->bias[0] + set * 32;
#pragma unroll 1
            for (int t = 0; t < kT1; ++t) {
                MBAR_WAIT(smem_u32(&hdr->acc1_full[ab][t]), ap);
                tc_fence_after();
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc1_col(ab, t) + uint32_t(set * 32) + lane_sel;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                const bool keep = interior_of(qm + t * kTile + row);
                tmem_ld_wait();
                uint8_t* dst = mid_smem + (set * 4) * kMidPitch + (t * kTile + row) * 16;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = keep ? fmaxf(__uint_as_float(v[pl * 8 + i]) + bs[pl * 8 + i], 0.f) : 0.f;
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
                mbar_arrive(smem_u32(&hdr->acc1_free[ab]));
                mbar_arrive(smem_u32(&hdr->mid_full));
            }
        };
        auto epilogue2 = [&](int j) {
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long q0 = unit * (kG * kTile) + row;
            // the residual = the block input at the output pixel: issued before the wait, it arrives under the MMAs
            uint4 res[kG][4];
            bool keep[kG];
#pragma unroll
            for (int g = 0; g < kG; ++g) {
                keep[g] = interior_of(q0 + g * kTile);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
                    res[g][pl] = keep[g] ? __ldg(reinterpret_cast<const uint4*>(p.in + ((long long)(set * 4 + pl) * p.in_plane + q0 + g * kTile) * 8))
                                         : make_uint4(0, 0, 0, 0);
            }
            MBAR_WAIT(smem_u32(&hdr->acc2_full), uint32_t(j & 1));
            tc_fence_after();
            const float* bs = hdr->bias[1] + set * 32;
#pragma unroll
            for (int g = 0; g < kG; ++g) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc2_col(g) + uint32_t(set * 32) + lane_sel;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                tmem_ld_wait();
                __half* outp = p.out + ((long long)(set * 4) * p.out_plane + q0 + g * kTile) * 8;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const uint32_t rr[4] = {res[g][pl].x, res[g][pl].y, res[g][pl].z, res[g][pl].w};
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r2 = unpack_half2(rr[i]);
                        f[2 * i] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i]) + bs[pl * 8 + 2 * i] + r2.x, 0.f);
                        f[2 * i + 1] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i + 1]) + bs[pl * 8 + 2 * i + 1] + r2.y, 0.f);
                    }
                    if (keep[g]) {
                        uint4 o4;
                        o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                        o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                        *reinterpret_cast<uint4*>(outp + (long long)pl * p.out_plane * 8) = o4;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&hdr->acc2_free));
        };
        // pipe order C1(j) C2(j) C1(j + 1) ...: E1(j) converts tile by tile under C1(j)'s later tiles, E2(j) runs under C1(j + 1)
        for (int j = 0; j < my_units; ++j) {
            epilogue1(j);
            epilogue2(j);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// shared memory with the minimum ring; *ring = the depth the 227 KB allow (a conv2 stage is 384 cycles of MMAs, a stage
// fetch out of L2 takes ~2000: four stages in flight throttle conv2, six or more do not)
size_t block64_smem(int W, int* in_pitch, int* run_bytes, int* ring) {
    const int R = W + 3;
    *run_bytes = (kT1 * kTile + 2 * R) * 16;
    *in_pitch = (*run_bytes + 127) / 128 * 128;
    const size_t base = size_t(kHeader) + kWBytes + size_t(kRing) * kTapBytes + size_t(kPlanes) * *in_pitch + kMidBytes;
    *ring = kRing;
    if (base <= 227 * 1024) *ring = std::min<int>(kMaxRing, kRing + int((227 * 1024 - base) / kTapBytes));
    return base;
}

}  // namespace

// 1 iff the streamed-W2 flavour takes the geometry (called by poco_basic_block_supported / _run in bblock_tc.cu)
int basic_block64_supported(int H, int W) {
    if (H < 1 || W < 1 || W + 3 > kLead) return 0;
    int pitch, run, ring;
    // both ends of every input run stay inside the activation guard: it starts kLead + R pixels before the unit's first
    // output pixel and ends 3 * 128 - kLead + R pixels after it
    return block64_smem(W, &pitch, &run, &ring) <= 227 * 1024 && (kT1 * kTile - kLead + W + 3) * 16 <= POCO_ACT_GUARD_BYTES;
}

int basic_block64_launch(const poco_basic_block* d, cudaStream_t stream) {
    const poco_act &in = d->in, &out = d->out;
    Block64Params p{};
    p.in = static_cast<const __half*>(in.data);
    p.out = static_cast<__half*>(out.data);
    p.w1 = static_cast<const __half*>(d->weight1);
    p.w2 = static_cast<const __half*>(d->weight2);
    p.b1 = d->bias1;
    p.b2 = d->bias2;
    p.in_plane = in.plane_stride;
    p.out_plane = out.plane_stride;
    p.H = in.H; p.W = in.W;
    p.P = int(int64_t(in.N) * (in.H + 2) * (in.W + 2));
    p.num_units = (p.P + kG * kTile - 1) / (kG * kTile);
    size_t smem = block64_smem(in.W, &p.in_pitch, &p.run_bytes, &p.ring);
    POCO_CHECK(smem <= 227 * 1024, "shared memory");
    smem += size_t(p.ring - kRing) * kTapBytes;
    static const char* prof_env = getenv("POCO_BBLOCK_PROF");       // bring-up: device address (decimal) of 16 zeroed uint64 counters
    p.prof = prof_env ? reinterpret_cast<unsigned long long*>(strtoull(prof_env, nullptr, 10)) : nullptr;
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(basic_block64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const int budget = d->max_ctas > 0 ? std::min(d->max_ctas, sm_count()) : sm_count();
    const int grid = std::max(1, std::min(p.num_units, budget));
    basic_block64_kernel<<<grid, kThreads, smem, stream>>>(p);
    POCO_LAUNCHED();
    return 0;
}

}  // namespace poco
