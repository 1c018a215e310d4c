// poco_b200 -- the BasicBlocks of a low-resolution HRNet branch as ONE tcgen05 launch with the crop RESIDENT in shared
// memory (sm_100a):
//     x <- ReLU(BN2(conv2(ReLU(BN1(conv1(x))))) + x)      n_blocks times     (hrnet.py:42-58 inside
//     HighResolutionModule._make_one_branch, hrnet.py:140-186)
// for the 128-channel 14x14 branch: 56 of the 311 conv launches of a POCO-CLIFF / HRNet-W32 forward.  A padded 16x16 crop
// is exactly two M = 128 tiles and its 128 channels are 64 KB of fp16, so the work unit is ONE CROP: its activations stay
// in shared memory for all 2 * n_blocks convs (there is no halo to recompute -- the whole image is on the SM), only the
// weights stream (16 KB half-tap stages through a five-deep ring, 288 KB per conv out of L2).  As separate poco_conv
// launches each of the eight convs of a branch pays the launch gap + pipeline fill for ~5 us of tensor work, writes its
// output to HBM and reads it back, and re-streams its weights per tile.
//
//   shared memory: X (block input / residual / block output, [plane][256 pixels][8 ch]) | MID (conv1's output, same
//   layout) | weight ring.  Both are in the planar operand layout, so the nine taps of a conv are nine shifted
//   descriptors into those bytes (common.cuh); what a shifted tile reads outside its plane (other planes, the bias
//   table, ring bytes) only reaches accumulator rows of halo pixels, and the epilogues never keep those.
//   conv1: A = X, epilogue 1 writes fp16 ReLU(acc + shift1) into MID (zero at halo pixels = conv2's padding);
//   conv2: A = MID, epilogue 2 adds shift2 and the residual read from X and writes the block output over X in place
//   (interior pixels only: the zero halo that came with the crop stays), or to global memory after the last block.
//
// Roles (352 threads, one persistent CTA per SM, crops dealt round-robin): warp 0 producer (the crop: one bulk copy per
// plane; then the weight stages of every conv, running ahead of the MMAs by the ring depth -- weights do not depend on
// activations, so the next conv's first stages are already there when its input is), warps 1 / 2 issue the MMAs of
// tile 0 / tile 1 (M = 128, N = 128: four K = 16 MMAs per stage and tile, 64 cycles each), warps 3-10 epilogue (four TMEM
// lane groups per tile).  Within a crop the convs are strictly sequential, so the tensor pipe idles while an epilogue
// runs (~3000 of 12700 cycles per conv, measured: two epilogue warps per scheduler cannot hide their tcgen05.ld / LDS
// latencies -- it is not the TMEM read rate, tools/tmem_ld_bench.cu); across CTAs nothing is shared but the weights in L2.
// mbarriers: w_full / w_empty[5] (producer <-> both issuers), x_full / x_free (crop landed / crop finished),
// acc_full[2] (issuer -> its tile's epilogue warps), tile_ready[2] (a tile's epilogue warps -> issuers: its rows of MID or X
// are written, its accumulator drained).  The two tiles share every weight stage but not their pace: the ring lets one
// issuer lead the other by up to five stages, and because a conv starts with the middle filter row (tap_of), which reads a
// tile's own rows only, a tile's next conv starts right after its OWN epilogue while the other tile's is still running.
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

constexpr int kC = 128;                     // channels (= N of every MMA)
constexpr int kPlanes = kC / 8;
constexpr int kTile = 128;
constexpr int kCropPix = 2 * kTile;         // pixel slots of a crop in shared memory (two tiles)
constexpr int kPitch = kCropPix * 16;       // bytes of one plane in shared memory
constexpr int kActBytes = kPlanes * kPitch; // 64 KB: one activation buffer
constexpr int kSlab = kC * 16;              // one (tap, 8-channel) weight slab: 128 output channels x 16 B
constexpr int kStageBytes = 8 * kSlab;      // half a tap: 8 planes = four K = 16 steps
constexpr int kStagesPerConv = 18;
constexpr int kRing = 5;
constexpr int kMaxConvs = 2 * POCO_MAX_BRANCH_BLOCKS;
constexpr int kHeader = 1024;
constexpr int kBiasBytes = kMaxConvs * kC * 4;
constexpr int kSmemBytes = kHeader + kBiasBytes + 2 * kActBytes + kRing * kStageBytes;
constexpr int kThreads = 352;               // producer, two issuers, eight epilogue warps
constexpr int kTmemCols = 256;              // two accumulators of 128 fp32 columns
static_assert(kSmemBytes <= 227 * 1024, "shared memory");

// Stage i of a conv holds half of filter tap tap_of(i >> 1): the middle filter row first.  Its taps read a tile's own image
// rows only (when 128 pixels are whole rows), so after an epilogue a tile's next conv can start before the OTHER tile's
// epilogue has finished; the rows above / below come six stages later.
__host__ __device__ constexpr int tap_of(int i) { return i < 3 ? i + 3 : (i < 6 ? i - 3 : i); }

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (one row per thread)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

struct BranchParams {
    const __half* in;
    __half* out;
    const __half* w[kMaxConvs];
    const float* b[kMaxConvs];
    long long in_plane, out_plane;          // plane strides in pixels
    int H, W;
    int n_crops;
    int n_convs;
    unsigned long long* prof;               // bring-up (POCO_BRANCH_PROF): [issuer 0: total, x_full, act_ready, w_full, issue, crops | epilogue warp 3:
                                            //  total, acc_full wait, work, - , - , convs]
};

struct Header {
    unsigned long long w_full[kRing], w_empty[kRing];
    unsigned long long x_full, x_free, acc_full[2], tile_ready[2];
    uint32_t tmem_base;
};
static_assert(sizeof(Header) <= kHeader, "header too large");

__global__ void __launch_bounds__(kThreads, 1) branch_kernel(const BranchParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Header* hdr = reinterpret_cast<Header*>(smem);
    float* bias_smem = reinterpret_cast<float*>(smem + kHeader);       // [conv][128]
    uint8_t* x_smem = smem + kHeader + kBiasBytes;
    uint8_t* mid_smem = x_smem + kActBytes;
    uint8_t* ring = mid_smem + kActBytes;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int Wp = p.W + 2, HpWp = (p.H + 2) * Wp;
    const int NC = p.n_convs;
    const int my_crops = (p.n_crops - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

    if (threadIdx.x < kRing) {
        mbar_init(smem_u32(&hdr->w_full[threadIdx.x]), 1);              // the producer's arrive.expect_tx
        mbar_init(smem_u32(&hdr->w_empty[threadIdx.x]), 2);             // one commit per issuer
    }
    if (threadIdx.x == 32) {
        mbar_init(smem_u32(&hdr->x_full), 1);
        mbar_init(smem_u32(&hdr->x_free), 8);                           // the eight epilogue warps
        mbar_init(smem_u32(&hdr->acc_full[0]), 1);
        mbar_init(smem_u32(&hdr->acc_full[1]), 1);
        mbar_init(smem_u32(&hdr->tile_ready[0]), 4);                    // the four epilogue warps of a tile
        mbar_init(smem_u32(&hdr->tile_ready[1]), 4);
    }
    for (int i = threadIdx.x; i < NC * kC; i += kThreads) bias_smem[i] = p.b[i / kC][i % kC];
    mbar_fence_init();
    if (warp == 1) tmem_alloc(smem_u32(&hdr->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    if (warp == 0) {
        // ============================================================ producer
        uint32_t slot = 0, par = 0;             // ring position of the next stage; par = parity of that slot's current use
        for (int j = 0; j < my_crops; ++j) {
            const long long crop = (long long)blockIdx.x + (long long)j * gridDim.x;
            const int x_at = j == 0 ? 0 : kRing;                        // (later crops: the first ring of stages goes out while
                                                                        //  the previous crop's last epilogue still reads X)
            for (int c = 0; c < NC; ++c) {
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w[c]);
                for (int s = 0; s < kStagesPerConv; ++s) {
                    if (c == 0 && s == x_at) {
                        MBAR_WAIT(smem_u32(&hdr->x_free), uint32_t(j & 1) ^ 1u);    // the previous crop is finished with X
                        if (elect_one()) {
                            const uint32_t bar = smem_u32(&hdr->x_full);
                            const uint32_t bytes = uint32_t(HpWp) * 16u;
                            mbar_arrive_expect_tx(bar, uint32_t(kPlanes) * bytes);
                            const __half* src = p.in + crop * HpWp * 8;
                            for (int pl = 0; pl < kPlanes; ++pl, src += p.in_plane * 8)
                                bulk_g2s(smem_u32(x_smem) + uint32_t(pl * kPitch), src, bytes, bar);
                        }
                        __syncwarp();
                    }
                    MBAR_WAIT(smem_u32(&hdr->w_empty[slot]), par ^ 1u);             // both issuers' MMAs on this slot retired
                    if (elect_one()) {
                        const uint32_t bar = smem_u32(&hdr->w_full[slot]);
                        mbar_arrive_expect_tx(bar, uint32_t(kStageBytes));
                        bulk_g2s(smem_u32(ring) + slot * uint32_t(kStageBytes), wsrc + size_t(tap_of(s >> 1) * 2 + (s & 1)) * kStageBytes,
                                 uint32_t(kStageBytes), bar);
                    }
                    __syncwarp();
                    if (++slot == kRing) { slot = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp <= 2) {
        // ============================================================ MMA issuers: warp 1 tile 0, warp 2 tile 1
        const uint32_t tile = uint32_t(warp - 1);
        const uint32_t idesc = umma_idesc_f16(kTile, kC);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);             // SBO = 128 B, descriptor version 1
        constexpr uint32_t a_lbo = (uint32_t(kPitch) >> 4) << 16, a_kstep = (2u * uint32_t(kPitch)) >> 4;
        constexpr uint32_t b_lbo = (uint32_t(kSlab) >> 4) << 16, b_kstep = (2u * uint32_t(kSlab)) >> 4;
        const uint32_t d_tmem = tmem_base + tile * uint32_t(kC);
        const uint32_t x16 = (smem_u32(x_smem) >> 4) + tile * uint32_t(kTile);
        const uint32_t mid16 = (smem_u32(mid_smem) >> 4) + tile * uint32_t(kTile);
        const uint32_t ring16 = smem_u32(ring) >> 4;
        uint32_t slot = 0, par = 0, ready_par = 0;
        const bool rows_aligned = (kTile % Wp) == 0;    // a tile is whole image rows: the middle filter row never leaves it
        const bool prof = p.prof != nullptr && warp == 1;
        long long pt[4] = {0, 0, 0, 0}, pt_mark = prof ? clock64() : 0;
        const long long pt_t0 = pt_mark;
        auto lap = [&](int k) { if (prof) { const long long t = clock64(); pt[k] += t - pt_mark; pt_mark = t; } };
        for (int j = 0; j < my_crops; ++j) {
            MBAR_WAIT(smem_u32(&hdr->x_full), uint32_t(j & 1));
            lap(0);
            for (int c = 0; c < NC; ++c) {
                if (c > 0) {
                    MBAR_WAIT(smem_u32(&hdr->tile_ready[tile]), ready_par);         // this tile's rows of the previous conv's output are in
                    if (!rows_aligned) MBAR_WAIT(smem_u32(&hdr->tile_ready[tile ^ 1u]), ready_par);     // shared memory, its accumulator drained
                    lap(1);
                }
                tc_fence_after();
                const uint32_t src16 = (c & 1) ? mid16 : x16;
#pragma unroll
                for (int s = 0; s < kStagesPerConv; ++s) {
                    const int tap = tap_of(s >> 1), half = s & 1;
                    const uint32_t shift = uint32_t((tap / 3 - 1) * Wp + (tap % 3 - 1));            // pixels = 16-byte units
                    if (s == 6 && c > 0 && rows_aligned) {      // the rows above / below reach into the other tile
                        MBAR_WAIT(smem_u32(&hdr->tile_ready[tile ^ 1u]), ready_par);
                        tc_fence_after();
                        lap(1);
                    }
                    if (s == kStagesPerConv - 1 && c > 0) ready_par ^= 1u;
                    MBAR_WAIT(smem_u32(&hdr->w_full[slot]), par);
                    lap(2);
                    if (elect_one()) {
                        const uint32_t a0 = ((src16 + shift + uint32_t(half * 8) * (uint32_t(kPitch) >> 4)) & 0x3FFFu) | a_lbo;
                        const uint32_t b0 = (ring16 + slot * (uint32_t(kStageBytes) >> 4)) | b_lbo;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(d_tmem, desc64(desc_hi, a0 + uint32_t(k) * a_kstep), desc64(desc_hi, b0 + uint32_t(k) * b_kstep), idesc,
                                     (s | k) ? 1u : 0u);
                        umma_commit(smem_u32(&hdr->w_empty[slot]));
                        if (s == kStagesPerConv - 1) umma_commit(smem_u32(&hdr->acc_full[tile]));
                    }
                    __syncwarp();
                    lap(3);
                    if (++slot == kRing) { slot = 0; par ^= 1u; }
                }
            }
        }
        if (prof && lane == 0) {
            atomicAdd(p.prof + 0, (unsigned long long)(clock64() - pt_t0));
            for (int k = 0; k < 4; ++k) atomicAdd(p.prof + 1 + k, (unsigned long long)pt[k]);
            atomicAdd(p.prof + 5, (unsigned long long)my_crops);
        }
    } else {
        // ============================================================ epilogue (8 warps): warp -> (tile, TMEM lane group)
        const int tile = (warp - 3) >> 2;
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int q = tile * kTile + lg * 32 + lane;    // pixel slot of this thread's accumulator row
        const uint32_t taddr0 = tmem_base + uint32_t(tile * kC) + (uint32_t(lg * 32) << 16);
        const int yy = q / Wp, xx = q - yy * Wp;
        const bool keep = q < HpWp && yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W;
        uint32_t acc_par = 0;
        const bool prof = p.prof != nullptr && warp == 3;
        long long pe[2] = {0, 0}, pe_mark = prof ? clock64() : 0;
        const long long pe_t0 = pe_mark;
        auto elap = [&](int k) { if (prof) { const long long t = clock64(); pe[k] += t - pe_mark; pe_mark = t; } };
        for (int j = 0; j < my_crops; ++j) {
            const long long crop = (long long)blockIdx.x + (long long)j * gridDim.x;
            MBAR_WAIT(smem_u32(&hdr->x_full), uint32_t(j & 1));         // (the residual reads below see the landed crop)
            for (int c = 0; c < NC; ++c) {
                const bool second = (c & 1) != 0, last = c == NC - 1;
                MBAR_WAIT(smem_u32(&hdr->acc_full[tile]), acc_par);
                acc_par ^= 1u;
                elap(0);
                tc_fence_after();
                const float* bs = bias_smem + c * kC;
                // 32 accumulator columns (4 planes) per step; the TMEM load of the next step is in flight while this one is processed
                auto process = [&](const uint32_t* v, int ch) {
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
                        const int plane = ch * 4 + pl;
                        const float4 ba = *reinterpret_cast<const float4*>(bs + plane * 8), bb = *reinterpret_cast<const float4*>(bs + plane * 8 + 4);
                        float f[8] = {__uint_as_float(v[pl * 8 + 0]) + ba.x, __uint_as_float(v[pl * 8 + 1]) + ba.y,
                                      __uint_as_float(v[pl * 8 + 2]) + ba.z, __uint_as_float(v[pl * 8 + 3]) + ba.w,
                                      __uint_as_float(v[pl * 8 + 4]) + bb.x, __uint_as_float(v[pl * 8 + 5]) + bb.y,
                                      __uint_as_float(v[pl * 8 + 6]) + bb.z, __uint_as_float(v[pl * 8 + 7]) + bb.w};
                        uint4* xq = reinterpret_cast<uint4*>(x_smem + plane * kPitch + q * 16);
                        if (second) {
                            const uint4 r4 = *xq;
                            const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 r2 = unpack_half2(rr[i]);
                                f[2 * i] += r2.x;
                                f[2 * i + 1] += r2.y;
                            }
                        }
                        uint4 o4;
                        o4.x = pack_half2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f)); o4.y = pack_half2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f));
                        o4.z = pack_half2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f)); o4.w = pack_half2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f));
                        if (!second) {
                            *reinterpret_cast<uint4*>(mid_smem + plane * kPitch + q * 16) = keep ? o4 : make_uint4(0, 0, 0, 0);
                        } else if (keep) {
                            if (!last)
                                *xq = o4;
                            else
                                *reinterpret_cast<uint4*>(p.out + ((long long)plane * p.out_plane + crop * HpWp + q) * 8) = o4;
                        }
                    }
                };
                uint32_t va[32], vb[32];
                tmem_ld32(taddr0, va);
#pragma unroll
                for (int ch = 0; ch < kC / 32; ++ch) {
                    uint32_t* cur = (ch & 1) ? vb : va;
                    uint32_t* nxt = (ch & 1) ? va : vb;
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(cur[i]));      // (uses of cur stay behind the wait)
                    if (ch + 1 < kC / 32) tmem_ld32(taddr0 + uint32_t((ch + 1) * 32), nxt);
                    process(cur, ch);
                }
                tc_fence_before();
                fence_proxy_async_smem();       // generic-proxy stores -> tcgen05.mma operand reads / the next crop's bulk copy
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(last ? &hdr->x_free : &hdr->tile_ready[tile]));
                elap(1);
            }
        }
        if (prof && lane == 0) {
            atomicAdd(p.prof + 8, (unsigned long long)(clock64() - pe_t0));
            atomicAdd(p.prof + 9, (unsigned long long)pe[0]);
            atomicAdd(p.prof + 10, (unsigned long long)pe[1]);
            atomicAdd(p.prof + 13, (unsigned long long)(my_crops * NC));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

}  // namespace poco

using namespace poco;

extern "C" int poco_branch_supported(int32_t C, int32_t H, int32_t W, int32_t n_blocks) {
    // one padded crop = at most two 128-pixel tiles; 128 channels (two 64 KB activation buffers + the weight ring)
    return C == kC && H >= 1 && W >= 1 && (H + 2) * (W + 2) <= kCropPix && n_blocks >= 1 && n_blocks <= POCO_MAX_BRANCH_BLOCKS;
}

extern "C" int poco_branch_run(const poco_branch* d, void* stream) {
    POCO_CHECK(d != nullptr, "null descriptor");
    if (check_act(d->in, "in") || check_act(d->out, "out")) return 1;
    const poco_act &in = d->in, &out = d->out;
    POCO_CHECK(in.C == out.C && in.N == out.N && in.H == out.H && in.W == out.W, "branch: in and out must share one geometry");
    POCO_CHECK(poco_branch_supported(in.C, in.H, in.W, d->n_blocks), "branch: only 128 channels, (H + 2) * (W + 2) <= 256, 1..4 blocks");
    POCO_CHECK(in.lo == nullptr && out.lo == nullptr, "branch: fp16 mode only");
    BranchParams p{};
    p.n_convs = 2 * d->n_blocks;
    for (int i = 0; i < p.n_convs; ++i) {
        POCO_CHECK(d->weight[i] != nullptr && d->bias[i] != nullptr, "null weight / bias");
        p.w[i] = static_cast<const __half*>(d->weight[i]);
        p.b[i] = d->bias[i];
    }
    p.in = static_cast<const __half*>(in.data);
    p.out = static_cast<__half*>(out.data);     // (may alias `in`: a CTA has read its whole crop before it writes it)
    p.in_plane = in.plane_stride;
    p.out_plane = out.plane_stride;
    p.H = in.H; p.W = in.W;
    p.n_crops = in.N;
    static const char* prof_env = getenv("POCO_BRANCH_PROF");        // bring-up: device address (decimal) of 16 zeroed uint64 counters
    p.prof = prof_env ? reinterpret_cast<unsigned long long*>(strtoull(prof_env, nullptr, 10)) : nullptr;
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); });
    const int budget = d->max_ctas > 0 ? std::min(d->max_ctas, sm_count()) : sm_count();
    const int grid = std::max(1, std::min(p.n_crops, budget));
    branch_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
    POCO_LAUNCHED();
    return 0;
}
