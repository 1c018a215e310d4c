// poco_b200 -- one residual BasicBlock as ONE tcgen05 launch (sm_100a):
//     out = ReLU(BN2(conv2(ReLU(BN1(conv1(in))))) + in)          (hrnet.py:42-58, downsample is None)
// for the 32-channel 56x56 and the 64-channel 28x28 branches of the HRNet modules (hrnet.py:140-186): 128 of the 311
// conv launches of a POCO-CLIFF / HRNet-W32 forward.  (Described for C = 32; the 64-channel flavour is at the end.)  As two poco_conv launches a block pays the fixed launch gap + pipeline fill twice,
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
//
// C = 64 (template flavour <64, 1, 1>): both weight tensors take 144 KB, so a unit is ONE conv2 tile (G = 1, two conv1
// tiles: 1.5x the MMAs of the two-launch block) and the input run and the intermediate have ONE shared-memory buffer each
// (the accumulators stay double buffered).  The tensor pipe then idles only while epilogue 1 runs between C1(u) and
// C2(u): the next input run lands under C2(u), C1(u + 1) follows C2(u) directly.  Measured at batch 256, 28x28: 61.8 us
// against 58.7 us for the two launches (6400-6700 cycles per unit for 5184 cycles of MMAs; even at the pipe's rate the
// recompute leaves ~5 us).  The plan's 64-channel blocks run on bblock64_tc.cu instead (conv2's weights streamed, two conv2
// tiles per unit); this flavour remains behind POCO_B200_BLOCK64_RESIDENT=1 and for rows too long for that kernel.
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

constexpr int kTile = 128;
constexpr int kLead = 64;               // conv1 starts this many pixels before the unit's first output pixel (>= W + 3)
constexpr int kHeader = 1024;
constexpr int kThreads = 352;           // producer, two issuers, eight epilogue warps
constexpr int kTmemCols = 512;

// C channels (= N of every MMA), G conv2 tiles per unit (G + 1 conv1 tiles), NBUF shared-memory buffers for the input
// run and for the intermediate
template <int C, int G, int NBUF>
struct Cfg {
    static constexpr int kPlanes = C / 8;
    static constexpr int kT1 = G + 1;
    static constexpr int kSlab = C * 16;                    // one (tap, 8-channel) weight slab: C output channels x 16 B
    static constexpr int kWBytes = 9 * kPlanes * kSlab;     // one conv's weights
    static constexpr int kMidPitch = kT1 * kTile * 16;      // one plane of the intermediate
    static constexpr int kChunks = C / 32;                  // epilogue work items (32 accumulator columns) per tile
    static_assert(2 * (kT1 + G) * C <= kTmemCols, "accumulators do not fit in TMEM");
};

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
    int run_bytes;                      // bytes of one input run: ((G + 1) * 128 + 2 R) * 16
    unsigned long long* prof;           // bring-up (POCO_BBLOCK_PROF): per issuer [total, in_full, acc1_free, mid_full, acc2_free, issue, units, ctas]
    __half* s2d_out;                    // second output: phase-split copy of `out` (poco_basic_block.out_s2d), or nullptr
    long long s2d_plane;                // its plane stride (pixels)
    int s2d_Wp, s2d_HpWp;               // its padded row pitch / padded pixels per crop
};

struct Header {
    unsigned long long in_full[2], in_free[2], acc1_full[2], acc1_free[2], mid_full[2], mid_free[2], acc2_full[2], acc2_free[2];
    unsigned long long w_full;
    uint32_t tmem_base;
    uint32_t pad_;
    float bias[2][64];
};
static_assert(sizeof(Header) <= kHeader, "header too large");

template <int C, int G, int NBUF>
__global__ void __launch_bounds__(kThreads, 1) basic_block_kernel(const BlockParams p) {
    using K = Cfg<C, G, NBUF>;
    constexpr int kPlanes = K::kPlanes, kT1 = K::kT1, kWBytes = K::kWBytes, kMidPitch = K::kMidPitch, kChunks = K::kChunks;
    extern __shared__ __align__(1024) uint8_t smem[];
    Header* hdr = reinterpret_cast<Header*>(smem);
    uint8_t* w_smem = smem + kHeader;                               // [conv][tap][plane][C][8]
    uint8_t* in_smem = w_smem + 2 * kWBytes;                        // [NBUF][plane][run]
    const int in_buf_bytes = kPlanes * p.in_pitch;
    uint8_t* mid_smem = in_smem + NBUF * in_buf_bytes;              // [NBUF][plane][(G + 1) * 128 pixels]
    constexpr int mid_buf_bytes = kPlanes * kMidPitch;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int Wp = p.W + 2, HpWp = (p.H + 2) * Wp, R = Wp + 1;
    const int my_units = (p.num_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    // ring positions of this CTA's j-th unit: shared-memory buffers (NBUF deep) and accumulators (two deep)
    auto sbuf = [](int j) { return NBUF == 2 ? uint32_t(j) & 1u : 0u; };
    auto spar = [](int j) { return (NBUF == 2 ? uint32_t(j) >> 1 : uint32_t(j)) & 1u; };
    auto abuf = [](int j) { return uint32_t(j) & 1u; };
    auto apar = [](int j) { return (uint32_t(j) >> 1) & 1u; };

    if (threadIdx.x < 16) {
        unsigned long long* bars = hdr->in_full;                    // the sixteen ring barriers are contiguous
        const int which = threadIdx.x >> 1;                         // 0 in_full 1 in_free 2 acc1_full 3 acc1_free 4 mid_full 5 mid_free 6 acc2_full 7 acc2_free
        const uint32_t count = (which == 3 || which == 4 || which == 7) ? 8u : 1u;      // the eight epilogue warps / one commit or producer
        mbar_init(smem_u32(bars + threadIdx.x), count);
    }
    if (threadIdx.x == 16) mbar_init(smem_u32(&hdr->w_full), 1);
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * C) {
        const int i = threadIdx.x - 64;
        hdr->bias[i / C][i % C] = (i < C ? p.b1 : p.b2)[i % C];
    }
    mbar_fence_init();
    if (warp == 1) tmem_alloc(smem_u32(&hdr->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    auto acc1_col = [&](uint32_t b, int t) { return (b * kT1 + uint32_t(t)) * uint32_t(C); };
    auto acc2_col = [&](uint32_t b, int g) { return (2u * kT1 + b * G + uint32_t(g)) * uint32_t(C); };

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
            const uint32_t b = sbuf(j), par = spar(j);
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            MBAR_WAIT(smem_u32(&hdr->in_free[b]), par ^ 1u);        // C1 of the unit that last used this buffer has retired
            if (elect_one()) {
                const uint32_t bar = smem_u32(&hdr->in_full[b]);
                mbar_arrive_expect_tx(bar, uint32_t(kPlanes) * uint32_t(p.run_bytes));
                const long long q0 = unit * (G * kTile) - kLead - R;        // (>= -8 KB guard, see the launcher)
                const __half* src = p.in + q0 * 8;
                const uint32_t dst = smem_u32(in_smem) + b * uint32_t(in_buf_bytes);
                for (int pl = 0; pl < kPlanes; ++pl, src += p.in_plane * 8)
                    bulk_g2s(dst + uint32_t(pl * p.in_pitch), src, uint32_t(p.run_bytes), bar);
            }
            __syncwarp();
        }
    } else if (warp <= 2) {
        // ============================================================ MMA issuers: warp 1 conv1, warp 2 conv2
        const uint32_t idesc = umma_idesc_f16(kTile, C);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t in_lbo = (uint32_t(p.in_pitch) >> 4) << 16, in_kstep = (2u * uint32_t(p.in_pitch)) >> 4;
        constexpr uint32_t mid_lbo = (uint32_t(kMidPitch) >> 4) << 16, mid_kstep = (2u * uint32_t(kMidPitch)) >> 4;
        constexpr uint32_t b_lbo = (uint32_t(K::kSlab) >> 4) << 16, b_kstep = (2u * K::kSlab) >> 4, b_tap = (uint32_t(kPlanes) * K::kSlab) >> 4;
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
            const uint32_t sb = sbuf(j), ab = abuf(j);
            MBAR_WAIT(smem_u32(&hdr->in_full[sb]), spar(j));
            lap(0);
            MBAR_WAIT(smem_u32(&hdr->acc1_free[ab]), apar(j) ^ 1u);
            lap(1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(in_smem) + sb * uint32_t(in_buf_bytes) + uint32_t(R) * 16u;
#pragma unroll
                for (int t = 0; t < kT1; ++t)
                    issue_linear<9, C / 16>(tmem_base + acc1_col(ab, t), ((a0 + uint32_t(t) * (kTile * 16u)) >> 4) | in_lbo, w1_lo, sh,
                                            in_kstep, b_kstep, b_tap, desc_hi, idesc, 0u);
                umma_commit(smem_u32(&hdr->acc1_full[ab]));
                umma_commit(smem_u32(&hdr->in_free[sb]));
            }
            __syncwarp();
            lap(4);
        };
        auto conv2 = [&](int j) {
            const uint32_t sb = sbuf(j), ab = abuf(j);
            MBAR_WAIT(smem_u32(&hdr->mid_full[sb]), spar(j));
            lap(2);
            MBAR_WAIT(smem_u32(&hdr->acc2_free[ab]), apar(j) ^ 1u);
            lap(3);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(mid_smem) + sb * uint32_t(mid_buf_bytes) + uint32_t(kLead) * 16u;
#pragma unroll
                for (int g = 0; g < G; ++g)
                    issue_linear<9, C / 16>(tmem_base + acc2_col(ab, g), ((a0 + uint32_t(g) * (kTile * 16u)) >> 4) | mid_lbo, w2_lo, sh,
                                            mid_kstep, b_kstep, b_tap, desc_hi, idesc, 0u);
                umma_commit(smem_u32(&hdr->acc2_full[ab]));
                umma_commit(smem_u32(&hdr->mid_free[sb]));
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
        // Work item = 32 accumulator columns (4 planes) of one tile; the items of a unit are dealt alternately to the two sets.
        const int ew = warp - 3;
        const int set = ew >> 2;
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
            const uint32_t sb = sbuf(j), ab = abuf(j);
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long qm = unit * (G * kTile) - kLead;                // pixel of row 0 of conv1 tile 0
            MBAR_WAIT(smem_u32(&hdr->acc1_full[ab]), apar(j));
            MBAR_WAIT(smem_u32(&hdr->mid_free[sb]), spar(j) ^ 1u);          // C2 of the unit that last read this buffer has retired
            tc_fence_after();
            uint8_t* mid = mid_smem + sb * mid_buf_bytes;
            for (int it = set; it < kT1 * kChunks; it += 2) {
                const int t = it / kChunks, ch = it % kChunks;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc1_col(ab, t) + uint32_t(ch * 32) + lane_sel;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                const bool keep = interior_of(qm + t * kTile + row);
                tmem_ld_wait();
                uint8_t* dst = mid + (ch * 4) * kMidPitch + (t * kTile + row) * 16;
                const float* bs = hdr->bias[0] + ch * 32;
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
                mbar_arrive(smem_u32(&hdr->mid_full[sb]));
            }
        };
        auto epilogue2 = [&](int j) {
            const uint32_t ab = abuf(j);
            const long long unit = (long long)blockIdx.x + (long long)j * gridDim.x;
            const long long q0 = unit * (G * kTile) + row;
            constexpr int kItems = G * kChunks;                             // 3 (C = 32, G = 3) or 2 (C = 64, G = 1)
            const int i0 = (kItems & 1) ? ((set + j) & 1) : set;            // an odd count: the set with the extra item alternates per unit
            // the residual = the block input at the output pixel: issued before the wait, it arrives under the MMAs
            uint4 res[2][4];
            bool keep[2];
            long long s2d_off[2] = {0, 0};      // this row's pixel in the phase-split second output (plane 0 of its phase)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int it = i0 + 2 * k, g = it / kChunks, ch = it % kChunks;
                keep[k] = it < kItems && interior_of(q0 + g * kTile);
                if (p.s2d_out != nullptr && keep[k]) {
                    const uint32_t q = uint32_t(q0 + g * kTile), n_crop = q / uint32_t(HpWp), rem = q - n_crop * uint32_t(HpWp);
                    const uint32_t y = rem / uint32_t(Wp) - 1u, x = rem % uint32_t(Wp) - 1u;
                    s2d_off[k] = ((long long)(((y & 1u) * 2u + (x & 1u)) * uint32_t(kPlanes)) * p.s2d_plane + (long long)n_crop * p.s2d_HpWp +
                                  ((y >> 1) + 1u) * uint32_t(p.s2d_Wp) + (x >> 1) + 1u) * 8;
                }
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
                    res[k][pl] = keep[k] ? __ldg(reinterpret_cast<const uint4*>(p.in + ((long long)(ch * 4 + pl) * p.in_plane + q0 + g * kTile) * 8))
                                         : make_uint4(0, 0, 0, 0);
            }
            MBAR_WAIT(smem_u32(&hdr->acc2_full[ab]), apar(j));
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int it = i0 + 2 * k, g = it / kChunks, ch = it % kChunks;
                if (it >= kItems) break;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc2_col(ab, g) + uint32_t(ch * 32) + lane_sel;
                tmem_ld16(taddr, v);
                tmem_ld16(taddr + 16, v + 16);
                tmem_ld_wait();
                __half* outp = p.out + ((long long)(ch * 4) * p.out_plane + q0 + g * kTile) * 8;
                const float* bs = hdr->bias[1] + ch * 32;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const uint32_t rr[4] = {res[k][pl].x, res[k][pl].y, res[k][pl].z, res[k][pl].w};
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r2 = unpack_half2(rr[i]);
                        f[2 * i] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i]) + bs[pl * 8 + 2 * i] + r2.x, 0.f);
                        f[2 * i + 1] = fmaxf(__uint_as_float(v[pl * 8 + 2 * i + 1]) + bs[pl * 8 + 2 * i + 1] + r2.y, 0.f);
                    }
                    if (keep[k]) {
                        uint4 o4;
                        o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                        o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                        *reinterpret_cast<uint4*>(outp + (long long)pl * p.out_plane * 8) = o4;
                        if (p.s2d_out != nullptr)       // the same pixel in the phase-split copy (feeds the stride-2 fuse convs)
                            *reinterpret_cast<uint4*>(p.s2d_out + s2d_off[k] + (long long)(ch * 4 + pl) * p.s2d_plane * 8) = o4;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&hdr->acc2_free[ab]));
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

template <int C, int G, int NBUF>
int launch_block(const poco_basic_block* d, BlockParams p, cudaStream_t stream) {
    using K = Cfg<C, G, NBUF>;
    const int R = d->in.W + 3;
    p.num_units = (p.P + G * kTile - 1) / (G * kTile);
    p.run_bytes = (K::kT1 * kTile + 2 * R) * 16;
    p.in_pitch = (p.run_bytes + 127) / 128 * 128;
    const size_t smem = size_t(kHeader) + 2 * K::kWBytes + size_t(NBUF) * K::kPlanes * p.in_pitch + size_t(NBUF) * K::kPlanes * K::kMidPitch;
    POCO_CHECK(smem <= 227 * 1024, "shared memory");
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(basic_block_kernel<C, G, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    const int budget = d->max_ctas > 0 ? std::min(d->max_ctas, sm_count()) : sm_count();
    const int grid = std::max(1, std::min(p.num_units, budget));
    basic_block_kernel<C, G, NBUF><<<grid, kThreads, smem, stream>>>(p);
    POCO_LAUNCHED();
    return 0;
}

}  // namespace

}  // namespace poco

using namespace poco;

extern "C" int poco_basic_block_supported(int32_t C, int32_t H, int32_t W) {
    // A unit's input run starts kLead + R pixels before its first output pixel and ends (G + 1) * 128 - kLead + R pixels
    // after it, R = W + 3: both ends of the tensor must stay inside the activation guard.  C = 64: the two weight tensors
    // (144 KB) leave room for one input run and one intermediate of two tiles only if the rows are short.
    if (H < 1 || W < 1 || W + 3 > kLead) return 0;
    if (C == 32) return (4 * kTile - kLead + W + 3) * 16 <= POCO_ACT_GUARD_BYTES;
    if (C == 64) {
        if (basic_block64_supported(H, W)) return 1;        // (bblock64_tc.cu)
        const int in_pitch = ((2 * kTile + 2 * (W + 3)) * 16 + 127) / 128 * 128;
        return kHeader + 2 * Cfg<64, 1, 1>::kWBytes + 8 * in_pitch + 8 * Cfg<64, 1, 1>::kMidPitch <= 227 * 1024;
    }
    return 0;
}

extern "C" int poco_basic_block_run(const poco_basic_block* d, void* stream) {
    POCO_CHECK(d != nullptr, "null descriptor");
    if (check_act(d->in, "in") || check_act(d->out, "out")) return 1;
    const poco_act &in = d->in, &out = d->out;
    POCO_CHECK(in.C == out.C && in.N == out.N && in.H == out.H && in.W == out.W, "basic block: in and out must share one geometry");
    POCO_CHECK(poco_basic_block_supported(in.C, in.H, in.W), "basic block: only 32 channels (W + 3 <= 64) or 64 channels (short rows) run fused");
    POCO_CHECK(in.lo == nullptr && out.lo == nullptr, "basic block: fp16 mode only");
    POCO_CHECK(in.data != out.data, "basic block: in and out must not alias");
    if (d->out_s2d.data != nullptr) {
        const poco_act& s2 = d->out_s2d;
        if (check_act(s2, "out_s2d")) return 1;
        POCO_CHECK(in.C == 32 && out.H % 2 == 0 && out.W % 2 == 0 && s2.C == 4 * out.C && s2.N == out.N && s2.H == out.H / 2 &&
                       s2.W == out.W / 2 && s2.lo == nullptr,
                   "basic block out_s2d: 32 channels only; 4 C channels at half the (even) output resolution");
    }
    POCO_CHECK(d->weight1 && d->weight2 && d->bias1 && d->bias2, "null weight / bias");
    const int64_t P = int64_t(in.N) * (in.H + 2) * (in.W + 2);
    POCO_CHECK(P + 4096 < (int64_t(1) << 31), "tensor too large");
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
    p.s2d_out = static_cast<__half*>(d->out_s2d.data);
    p.s2d_plane = d->out_s2d.plane_stride;
    p.s2d_Wp = d->out_s2d.W + 2;
    p.s2d_HpWp = (d->out_s2d.H + 2) * (d->out_s2d.W + 2);
    static const char* prof_env = getenv("POCO_BBLOCK_PROF");       // bring-up: device address (decimal) of 16 zeroed uint64 counters
    p.prof = prof_env ? reinterpret_cast<unsigned long long*>(strtoull(prof_env, nullptr, 10)) : nullptr;
    // 64 channels: the flavour with conv2's weights streamed (bblock64_tc.cu: two conv2 tiles per unit); POCO_B200_BLOCK64_RESIDENT=1
    // keeps both weight tensors resident and runs one conv2 tile per unit (slower, see the header)
    static const bool resident64 = getenv("POCO_B200_BLOCK64_RESIDENT") != nullptr;
    if (in.C == 64 && !resident64 && basic_block64_supported(in.H, in.W)) return basic_block64_launch(d, static_cast<cudaStream_t>(stream));
    if (in.C == 64) return launch_block<64, 1, 1>(d, p, static_cast<cudaStream_t>(stream));
    return launch_block<32, 3, 2>(d, p, static_cast<cudaStream_t>(stream));
}
