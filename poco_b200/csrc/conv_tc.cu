// poco_b200 -- conv + folded-BN + residual + ReLU as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces every nn.Conv2d -> nn.BatchNorm2d(eval) -> (+residual) -> nn.ReLU chain of the reference
// backbones and the PARE conv branches (hrnet.py:42-58, :79-99, :198-240, :345-384, :437-450;
// hrnet_cls.py:306-353; resnet.py:100-121, :201-217; pare_head.py:468-491).
//
// GEMM view:  D[M = output pixels, N = Cout] = A[M, K = taps*Cin] * W[N, K]^T, fp16 x fp16 -> fp32.
//   * M index = padded-linear pixel index q of the OUTPUT plane; one CTA tile = 128 consecutive q.
//     Halo / out-of-range rows are computed and discarded (MMA rows are independent).
//   * accumulators: TMEM, 128 lanes x n_tile fp32 columns, double buffered (epilogue of tile i
//     overlaps the MMAs of tile i+1); one elected thread issues tcgen05.mma (M=128, N=n_tile, K=16).
//   * operands: shared memory in the UMMA no-swizzle K-major canonical layout.  The planar-8
//     activation layout already *is* that layout (16 B per pixel per plane), so
//       MODE_LINEAR (3x3/s1/p1 and 1x1/s1): ONE bulk copy (TMA engine, UBLKCP) per 8-channel plane
//         brings the tile plus its (W+2)+1 pixel halo; the 9 taps are 9 shifted descriptors into the
//         same bytes -- the input is read from L2/HBM once, not 9 times;
//       MODE_GATHER (stride 2, 7x7, anything else): 128 producer threads cp.async one 16-byte pixel
//         each per plane and tap (zero-fill outside the image).
//     Weights [tap][Cin/8][Cout][8] are kept resident in shared memory for the whole persistent CTA
//     when they fit, else streamed per K-chunk with bulk copies.
//   * epilogue (4 warps): tcgen05.ld 16 columns at a time -> +bias (BN shift) -> +residual -> ReLU ->
//     fp16 -> 16-byte stores, 512 contiguous bytes per warp; halo pixels are never written.
//   * persistent grid: <= one CTA per SM, static round-robin over M tiles; mbarrier pipelines
//     smem full/empty (producer <-> MMA) and tmem full/empty (MMA <-> epilogue).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "internal.h"

namespace poco {

namespace {

constexpr int kMaxStages = 16;                     // operand stages over both rings (header barrier arrays)
constexpr int kMaxAccBufs = 8;                     // TMEM accumulator ring (512 columns / n_tile)
constexpr int kTileM = 128;
constexpr int kSmemBudget = 225 * 1024;   // of 227 KB usable per CTA
constexpr int kSmemBudgetHalf = 110 * 1024;       // two CTAs per SM (small-N layers, see conv_tc_launch_chain)
constexpr int kHeaderBytes = 8192;        // barriers + bias + dx-in-N exchange buffers
constexpr int kGatherLag = 2;
constexpr int kPlaneTile = kTileM * 16;            // one output plane of one tile: 2 KB, contiguous in HBM
constexpr int kOutRing = 0, kResRing = 4, kMaxResRing = 16;     // residual prefetch depth grows into spare smem
// residual ring: 8 warps x depth x item slice; depth 0 = register mode (POCO_B200_RES_RING=0), no ring at all
// (A register-prefetch variant of the residual path -- one coalesced 512-byte load per plane issued an item
// ahead, no ring -- was measured and removed: carrying both paths cost 30 registers and 5-12 % on every
// residual conv, and alone it was slower than the ring for 32->32 @56: 49 vs 45.6 us.)
static int res_ring_base(bool half) {
    static const int v = [] { const char* e = getenv("POCO_B200_RES_RING"); return e ? std::max(1, atoi(e)) : kResRing; }();
    return half ? std::min(v, 2) : v;
}
// a conv without a residual input has no ring at all: its shared memory goes to operand stages (bytes in flight)
static bool ring_always() { static const bool v = [] { const char* e = getenv("POCO_B200_RING_ALWAYS"); return e && e[0] == '1'; }(); return v; }
static int max_stages() { static const int v = [] { const char* e = getenv("POCO_B200_MAX_STAGES"); return e ? std::max(2, std::min(kMaxStages, atoi(e))) : kMaxStages; }(); return v; }
static int ring_bytes_for(int item_planes, bool half, int sp, bool any_res) {
    return (any_res || ring_always()) ? 2 * (kOutRing + res_ring_base(half)) * item_planes * kPlaneTile * sp : 0;
}

enum { MODE_LINEAR = 0, MODE_GATHER = 1 };

constexpr int kMaxSegs = POCO_MAX_CHAIN;

// one convolution of a chain: same geometry as every other segment, own tensors
struct ChainSeg {
    const __half* in;
    __half* out;
    const __half* res;
    const __half* w;
    const float* bias;
    int relu;
    int pad_;
    // split-precision mode (SPLIT instantiations, single segment): the rounding-residual tensors
    const __half* in_lo;
    __half* out_lo;
    const __half* res_lo;
};

struct ConvTcParams {
    ChainSeg seg[kMaxSegs];
    int n_segs;           // > 1: MODE_LINEAR only; segment s+1 reads what segment s wrote, ordered by per-tile flags
    int flag_expect;      // epilogue-warp arrivals that complete one tile of one segment
    int* flags;           // [n_segs - 1][num_m_tiles], zeroed before the launch
    long long in_plane;
    int Hin, Win;
    long long out_plane;
    int Hout, Wout;
    long long res_plane;
    int Cin, Cout;
    int kh, kw, stride, pad;
    int m_group;          // MODE_LINEAR: adjacent tiles fetched as one run (one halo per group) and walked per stage: 1, 2 or 4
    int taps;             // filter taps the MMA loop walks (kh*kw; 3 in dx-in-N mode: the rows of the 3x3)
    int tile_stride;      // output pixels between consecutive tiles (128; 126 in dx-in-N mode) ...
    int tile_origin;      // ... and the pixel of tile 0 / row 0 (0; -1 in dx-in-N mode)
    int n_out;            // output channels this CTA's epilogue writes (n_tile; Cout in dx-in-N mode)
    int epi_items;        // epilogue work items per tile
    int w_bufs;           // resident-weight buffers (2 = the next segment's weights load while this one computes)
    int kc, n_chunks;
    int n_tile;
    int w_resident;
    int stages;           // per ring
    int rings;            // 1, or 2 = one shared-memory stage ring per MMA issuer warp (tiles alternate)
    int num_m_tiles;
    long long P_out;
    int a_plane_bytes, a_copy_bytes, a_stage_bytes, w_stage_bytes, w_res_bytes;
    int halo;
    int tmem_cols, acc_bufs;
    int res_ring;         // residual prefetch ring depth per epilogue half
    int item_planes;      // epilogue work item = item_planes x 8 accumulator columns (2 or 4)
    int tap_group;        // gather mode: filter taps per stage
    // space-to-depth plumbing of the stride-2 convs (see poco_conv.in_s2d / out_s2d)
    int in_s2d;           // the input is the phase-split tensor: 4 phase blocks of Cin/8 planes at the OUTPUT resolution
    int halo_hi;          // pixels fetched behind the tile run (= halo, 0 for in_s2d: its taps reach back only)
    __half* s2d_out;      // second output: phase-split copy of `out` (4 Cout channels at half resolution), or nullptr
    __half* s2d_out_lo;
    long long s2d_plane;  // its plane stride (pixels)
    int s2d_Wp, s2d_HpWp; // its padded row pitch / padded pixels per crop
    int s2d_only;         // skip the normal store
    int debug;            // POCO_CONV_DEBUG bits (bring-up only): 1 skip epilogue work, 2 skip MMAs, 4 skip A loads,
                          // 8 skip output stores, 16 ignore the residual, 32 cycle accounting of the MMA issuers
    unsigned long long* prof;   // debug & 32: [2 issuers][8] cycle sums (POCO_CONV_PROF points the launcher at a buffer)
};

struct SmemHeader {
    unsigned long long full[kMaxStages];
    unsigned long long empty[kMaxStages];
    unsigned long long tmem_full[kMaxAccBufs];
    unsigned long long tmem_empty[kMaxAccBufs];
    unsigned long long w_ready[2];
    unsigned long long w_free[2];
    unsigned long long res_full[8 * kMaxResRing];      // [epilogue warp][slot]
    uint32_t tmem_base;
    uint32_t pad_[3];
    float bias[2][256];                                // double buffered across chain segments
    float xch[2][2][4][2][32];                         // dx-in-N: [half][buffer][warp][D0 of row 31 | D2 of row 0][channel]
    unsigned long long stamps[8];                      // debug & 64: globaltimer of CTA 0's phases
};
static_assert(sizeof(SmemHeader) <= kHeaderBytes, "header too large");

template <int MODE>
struct Roles {
    // MODE_LINEAR: warp 0 producer, warps 1-2 MMA issuers (alternate tiles), warps 3-10 epilogue
    // MODE_GATHER: warps 0-3 A producers, warps 4-5 MMA issuers, warp 6 W producer, warps 7-14 epilogue
    static constexpr int kProducerWarps = MODE == MODE_LINEAR ? 1 : 4;
    static constexpr int kMmaWarp = MODE == MODE_LINEAR ? 1 : 4;
    static constexpr int kWWarp = MODE == MODE_LINEAR ? 0 : 6;
    static constexpr int kEpiWarp0 = MODE == MODE_LINEAR ? 3 : 7;
    static constexpr int kThreads = (kEpiWarp0 + 8) * 32;
};


template <int TG>
__device__ __forceinline__ void issue_gather(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t b_tap,
                                             uint32_t desc_hi, uint32_t idesc, uint32_t acc0) {
#if POCO_ISSUE_BATCH > 0
    uint32_t al[TG], bl[TG];
#pragma unroll
    for (int tt = 0; tt < TG; ++tt) {
        al[tt] = a_lo + uint32_t(tt) * 256u;
        bl[tt] = b_lo + uint32_t(tt) * b_tap;
    }
#pragma unroll
    for (int tt = 0; tt < TG; ++tt) asm volatile("" : "+r"(al[tt]), "+r"(bl[tt]));
#pragma unroll
    for (int tt = 0; tt < TG; ++tt) umma_f16(d_tmem, desc64(desc_hi, al[tt]), desc64(desc_hi, bl[tt]), idesc, tt ? 1u : acc0);
#else
#pragma unroll
    for (int tt = 0; tt < TG; ++tt)      // one K=16 MMA per tap of the group; A taps are 4 KB (256 units) apart
        umma_f16(d_tmem, desc64(desc_hi, a_lo + uint32_t(tt) * 256u), desc64(desc_hi, b_lo + uint32_t(tt) * b_tap), idesc,
                 tt ? 1u : acc0);
#endif
}

// global-memory helpers of the chain protocol
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_gpu_add(int* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// order earlier generic-proxy accesses (here: an acquire of data other threads wrote with st.global)
// before later async-proxy accesses (bulk copies reading that data)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// DXN ("dx in N", 3x3 / stride 1 with Cout <= 64): the three horizontal taps of a filter row ride in the N
// dimension (N = 3 * Cout), so a tile costs 3 * Cin/16 MMAs instead of 9 * Cin/16 and its A operand is
// fetched from shared memory 3 times instead of 9 -- for Cout = 32 / 64 the tensor pipe is otherwise busy
// re-reading A (ncu: 58 % tensor-pipe-active for 32->32 with 12 % of the MMA rate used).  The accumulator
// then holds D_s[p] = sum_{r,c} in[p + (r-1) Wp][c] W[r][s][c] and the epilogue forms
// out[q] = D_0[q-1] + D_1[q] + D_2[q+1] with warp shuffles (+ a 2 x 32-float exchange between the four warps
// of a tile); tiles advance by 126 pixels so that rows 0 and 127 of a tile are never outputs.
// MINB = 2: the "half" configuration (two CTAs per SM), which needs the register cap of __launch_bounds__(.., 2);
// the one-CTA instantiations keep their registers (capping them cost the wide-N epilogue 20 %).
// SPLIT: split-precision ("parity") mode.  Every activation is a pair of fp16 tensors (hi, lo = x - hi) and the weights
// are [W_hi][W_lo]; a K chunk lands as kc/8 hi planes followed by kc/8 lo planes, the issuer runs the same straight-line
// MMA sequence three times into one accumulator -- (x_hi, W_hi), (x_lo, W_hi), (x_hi, W_lo) -- and the epilogue adds
// residual hi + lo in fp32 and stores hi = fp16(y), lo = fp16(y - hi).  Everything else (roles, barriers, rings) is shared.
// SPLIT == 2 ("N concatenation", weight format 2, Cout <= 64): the weights are ONE tensor [tap][Cin/8][2 Cout][8] whose slab
// rows are W_hi then W_lo.  One N = 2 Cout MMA forms x_hi W_hi | x_hi W_lo side by side (the A operand -- 4 KB per MMA,
// what bounds small-N MMAs -- is fetched once for both products), one N = Cout MMA adds x_lo W_hi to the first half, and
// the epilogue sums the two column groups in fp32: 88 instead of 120 tensor cycles per K step at Cout = 32 (112 / 144 at
// 64), and the 2^-11-sized x_hi W_lo term gets an accumulator of its own.
// S2DOUT: the epilogue also writes the phase-split copy of the output (poco_conv.out_s2d).  A template parameter, not a
// run-time branch: carrying the extra address arithmetic and predicates in every instantiation cost the epilogue-bound
// layers up to 27 % (64->256 1x1 + residual: 205 -> 261 us) when it was tried as a branch.
template <int MODE, int IPL, bool DXN, int MINB = 1, int SPLIT = 0, bool S2DOUT = false>
__global__ void __launch_bounds__(Roles<MODE>::kThreads, MINB) conv_tc_kernel(const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
    const bool tl_on = (p.debug & 64) && blockIdx.x == 0 && blockIdx.y == 0;     // timeline of CTA 0 (bring-up)
    if (tl_on && threadIdx.x == 0) hdr->stamps[0] = global_timer_ns();
    const int res_ring_n = p.res_ring;
    constexpr int ipl = IPL;                            // output planes (8 columns each) per epilogue work item
    constexpr int SP = SPLIT ? 2 : 1;
    constexpr bool NCAT = SPLIT == 2;
    const int warp_ring_bytes = (kOutRing + res_ring_n) * ipl * 512 * SP;      // per epilogue warp: out ring + residual ring
    uint8_t* w_res = smem + kHeaderBytes + 8 * warp_ring_bytes;
    uint8_t* stage0 = w_res + p.w_res_bytes * p.w_bufs;
    const int stage_bytes = p.a_stage_bytes + p.w_stage_bytes;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // provably warp-uniform for ptxas
    const int lane = threadIdx.x & 31;
    const int nb = blockIdx.y;                      // N block
    const int taps = p.taps;
    const int planes_per_chunk = p.kc >> 3;
    const int cin8 = p.Cin >> 3;
    const int kiters = MODE == MODE_LINEAR ? p.n_chunks : p.n_chunks * (taps / p.tap_group);
    const uint32_t slab_bytes = uint32_t(p.n_tile) * (NCAT ? 32u : 16u);     // one (tap, 8-channel) weight slab
    const int n_segs = MODE == MODE_LINEAR ? p.n_segs : 1;
    // Work unit = G adjacent tiles (G = 1 outside MODE_LINEAR): unit u of this CTA's j-th turn is
    // blockIdx.x + j * gridDim.x and covers tiles [u * G, u * G + G); only the globally last unit can be short.
    const int G = MODE == MODE_LINEAR ? p.m_group : 1;          // 1, 2 or 4
    const int g_log2 = G == 4 ? 2 : (G == 2 ? 1 : 0);
    const int num_units = (p.num_m_tiles + G - 1) >> g_log2;
    const int my_units = (num_units - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    auto unit_of = [&](int j) { return int(blockIdx.x) + j * int(gridDim.x); };
    auto unit_tiles = [&](int u) { return min(G, p.num_m_tiles - (u << g_log2)); };
    auto tile_of = [&](int jt) { return (unit_of(jt >> g_log2) << g_log2) + (jt & (G - 1)); };     // local tile index -> tile (no division)
    const int my_tiles = my_units > 0 ? ((my_units - 1) << g_log2) + unit_tiles(unit_of(my_units - 1)) : 0;   // per segment
    using R = Roles<MODE>;

    // ---------------------------------------------------------------- setup
    {   // barrier init spread over the CTA (one thread initialising ~170 mbarriers cost ~2.5 us per launch)
        const int t = threadIdx.x;
        const uint32_t full_count = MODE == MODE_LINEAR ? 1u : (128u + (p.w_resident ? 0u : 1u));
        if (t < p.stages * p.rings) {
            mbar_init(smem_u32(&hdr->full[t]), full_count);
            mbar_init(smem_u32(&hdr->empty[t]), 1);
        }
        if (t >= 32 && t < 32 + p.acc_bufs) {
            mbar_init(smem_u32(&hdr->tmem_full[t - 32]), 1);
            mbar_init(smem_u32(&hdr->tmem_empty[t - 32]), p.epi_items >= 2 ? 8 : 4);   // epilogue warps draining one tile
        }
        if (t >= 60 && t < 62) {
            mbar_init(smem_u32(&hdr->w_ready[t - 60]), 1);
            mbar_init(smem_u32(&hdr->w_free[t - 60]), uint32_t(p.rings));      // every MMA issuer commits once per segment
        }
        if (t >= 64 && t < 64 + 8 * kMaxResRing) mbar_init(smem_u32(&hdr->res_full[t - 64]), 1);
        mbar_fence_init();
    }
    if (warp == R::kMmaWarp) tmem_alloc(smem_u32(&hdr->tmem_base), uint32_t(p.tmem_cols));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    if (tl_on && threadIdx.x == 0) hdr->stamps[1] = global_timer_ns();
    // Programmatic dependent launch: let the next kernel of the stream start its prologue on every SM
    // this CTA leaves; roles that touch activations call pdl_wait() before their first access.
    pdl_launch_dependents();
    const uint32_t buf_cols = uint32_t(p.tmem_cols) / uint32_t(p.acc_bufs);
    const uint32_t nacc = uint32_t(p.acc_bufs);

    // ---------------------------------------------------------------- weight loader (shared by both modes)
    // All producer / MMA control flow below is warp-convergent with the single issuing lane chosen by
    // elect.sync: every operand of the bulk copies and of tcgen05.mma is then provably warp-uniform
    // and lives in uniform registers (a divergent `if (lane == 0)` loop made ptxas wrap each UBLKCP /
    // UTCHMMA in a vote + branch "waterfall", ~60 issue slots per MMA -- see profiles/).
    // When this CTA covers all of Cout (one N block), consecutive 8-channel slabs are contiguous in
    // global and shared memory, so `group` slabs travel as one bulk copy.
    const int w_n = DXN ? 3 * p.Cout : (NCAT ? 2 * p.Cout : p.Cout);       // columns of one (tap, 8-channel) slab of the weight tensor
    const bool whole_n = NCAT || p.n_tile == w_n;
    auto load_resident_weights = [&](int s) {      // weights of segment s -> resident buffer s % w_bufs
        const int b = p.w_bufs == 2 ? (s & 1) : 0;
        const __half* wg = p.seg[s].w + size_t(nb) * p.n_tile * 8;
        const uint32_t dst = smem_u32(w_res) + uint32_t(b) * uint32_t(p.w_res_bytes);
        const uint32_t bar = smem_u32(&hdr->w_ready[b]);
        mbar_arrive_expect_tx(bar, uint32_t(p.w_res_bytes));
        const int total = taps * cin8 * (NCAT ? 1 : SP);         // (split mode: the W_lo tensor follows W_hi, slab for slab; N concatenation: double-size slabs)
        const int group = whole_n ? min(total, 64) : 1;          // <= 64 slabs (<= 256 KB) per copy
        for (int i = 0; i < total; i += group) {
            const int g = min(group, total - i);
            bulk_g2s(dst + uint32_t(i) * slab_bytes, wg + size_t(i) * w_n * 8, uint32_t(g) * slab_bytes, bar);
        }
    };
    auto load_stage_weights = [&](const __half* wg, uint32_t ws, uint32_t bar, int t0, int t1, int c) {
        // taps [t0, t1) of K chunk c -> ws, slabs ordered [tap][plane]
        for (int t = t0; t < t1; ++t) {
            const uint32_t dst = ws + uint32_t((t - t0) * planes_per_chunk) * slab_bytes;
            const __half* src = wg + size_t(t * cin8 + c * planes_per_chunk) * w_n * 8;
            if (whole_n) {
                bulk_g2s(dst, src, uint32_t(planes_per_chunk) * slab_bytes, bar);
            } else {
                for (int j = 0; j < planes_per_chunk; ++j)
                    bulk_g2s(dst + uint32_t(j) * slab_bytes, src + size_t(j) * w_n * 8, slab_bytes, bar);
            }
        }
    };
    // Resident weights of a chain: segment s uses buffer s % w_bufs.  w_ready[b] completes once per
    // segment that uses buffer b; w_free[b] completes when the MMAs of such a segment have retired.
    const size_t w_lo_off = size_t(taps) * cin8 * w_n * 8;        // split mode: elements from W_hi to W_lo
    auto wbuf_of = [&](int s) { return p.w_bufs == 2 ? (s & 1) : 0; };
    auto wuse_of = [&](int s) { return p.w_bufs == 2 ? (s >> 1) : s; };      // how many earlier segments used that buffer

    if (MODE == MODE_LINEAR && warp == 0) {
        // ============================================================ producer (bulk copies)
        if (p.w_resident && elect_one()) load_resident_weights(0);      // weights are constants: no dependency
        __syncwarp();
        pdl_wait();
        // Ring positions are kept as (slot, phase) counters: `it % stages` / `it / stages` with a run-time divisor cost
        // ~150 cycles of dependent integer-division code per stage on every role's critical path (r02 ncu source page).
        uint32_t r_slot[2] = {0u, 0u}, r_ph[2] = {0u, 0u}, tl = 0, ul = 0;      // ul: units so far (the two issuers take alternate units)
        // double-buffered weights: fetch the next segment's a few tiles into this one (its buffer was
        // last read two segments ago); single buffer: after this segment's last MMA has retired
        const int w_prefetch_at = min(my_tiles - 1, 2 * p.stages * p.rings);
        for (int s = 0; s < n_segs; ++s) {
            const ChainSeg& sg = p.seg[s];
            const __half* wg = sg.w + size_t(nb) * p.n_tile * 8;
            const int* flags_prev = s > 0 ? p.flags + size_t(s - 1) * p.num_m_tiles : nullptr;
            int ready_upto = (flags_prev != nullptr && !(p.debug & 256)) ? 0 : my_tiles;     // local tiles [0, ready_upto) may be loaded
            for (int ju = 0; ju < my_units; ++ju) {
                const int unit = unit_of(ju), gcount = unit_tiles(unit);
                const int j = ju * G;                       // first local tile of the unit
                if (j + gcount > ready_upto) {
                    // Tile t reads output tiles t-1 .. t+1 of the previous segment.  One poll covers this
                    // CTA's next 10 tiles (3 flags each, one per lane): in steady state the neighbours
                    // finished them a whole segment ago, so the L2 round trip is paid once per 10 tiles.
                    uint32_t spins = 0;
                    unsigned long long t0 = 0;
                    while (j + gcount > ready_upto) {
                        const int jj = ready_upto + lane / 3;
                        const int tt = tile_of(jj) - 1 + lane % 3;
                        const bool need = lane < 30 && jj < my_tiles && tt >= 0 && tt < p.num_m_tiles;
                        const int v = need ? ld_acquire_gpu(flags_prev + tt) : p.flag_expect;
                        const uint32_t ok = __ballot_sync(0xffffffffu, v >= p.flag_expect);
                        int cnt = 0;
                        while (cnt < 10 && ((ok >> (3 * cnt)) & 7u) == 7u) ++cnt;
                        ready_upto = min(my_tiles, ready_upto + cnt);
                        if ((++spins & 255u) == 0u) {
                            if (t0 == 0) t0 = global_timer_ns();
                            spin_timeout(-(s * 100000 + j), t0);
                        }
                    }
                    fence_proxy_async_all();
                }
                // one contiguous run per plane: the unit's tiles plus ONE halo on each side
                const long long q0 = (long long)unit * G * p.tile_stride + p.tile_origin - p.halo;
                const uint32_t copy_bytes = uint32_t(((gcount - 1) * p.tile_stride + kTileM + p.halo + p.halo_hi) * 16);
                const uint32_t ring = p.rings == 2 ? (ul & 1u) : 0u;
                ++ul;
                for (int c = 0; c < p.n_chunks; ++c) {
                    const uint32_t slot = ring * p.stages + r_slot[ring], ph = r_ph[ring];
                    if (++r_slot[ring] == uint32_t(p.stages)) { r_slot[ring] = 0; r_ph[ring] ^= 1u; }
                    MBAR_WAIT(smem_u32(&hdr->empty[slot]), ph ^ 1u);
                    if (elect_one()) {
                        const uint32_t bar = smem_u32(&hdr->full[slot]);
                        const bool skip_a = (p.debug & 4) != 0;
                        const int phases = p.in_s2d ? 4 : 1;       // phase-split input: the chunk's channels of all four phases
                        const uint32_t tx = (skip_a ? 0u : uint32_t(planes_per_chunk * phases * SP) * copy_bytes) +
                                            (p.w_resident ? 0u : uint32_t(p.w_stage_bytes));
                        mbar_arrive_expect_tx(bar, tx);
                        const uint32_t st = smem_u32(stage0 + size_t(slot) * stage_bytes);
                        for (int ph = 0; ph < phases && !skip_a; ++ph) {
                            const long long src_off = ((long long)(ph * cin8 + c * planes_per_chunk) * p.in_plane + q0) * 8;
                            const __half* src = sg.in + src_off;
                            const uint32_t dst = st + uint32_t(ph * planes_per_chunk) * p.a_plane_bytes;
                            for (int jj = 0; jj < planes_per_chunk; ++jj, src += p.in_plane * 8)
                                bulk_g2s(dst + uint32_t(jj) * p.a_plane_bytes, src, copy_bytes, bar);
                            if (SPLIT) {        // the lo planes of the chunk land behind all of its hi planes
                                const __half* srl = sg.in_lo + src_off;
                                const uint32_t dstl = dst + uint32_t(phases * planes_per_chunk) * p.a_plane_bytes;
                                for (int jj = 0; jj < planes_per_chunk; ++jj, srl += p.in_plane * 8)
                                    bulk_g2s(dstl + uint32_t(jj) * p.a_plane_bytes, srl, copy_bytes, bar);
                            }
                        }
                        if (!p.w_resident) {
                            load_stage_weights(wg, st + p.a_stage_bytes, bar, 0, taps, c);
                            if (SPLIT == 1) load_stage_weights(wg + w_lo_off, st + p.a_stage_bytes + (p.w_stage_bytes >> 1), bar, 0, taps, c);
                        }
                    }
                    __syncwarp();
                }
                tl += uint32_t(gcount);
                if (p.w_resident && p.w_bufs == 2 && s + 1 < n_segs && j == w_prefetch_at) {
                    if (s >= 1) MBAR_WAIT(smem_u32(&hdr->w_free[wbuf_of(s + 1)]), uint32_t(wuse_of(s + 1) - 1) & 1u);
                    if (elect_one()) load_resident_weights(s + 1);
                    __syncwarp();
                }
            }
            if (p.w_resident && p.w_bufs == 1 && s + 1 < n_segs) {
                MBAR_WAIT(smem_u32(&hdr->w_free[0]), uint32_t(s) & 1u);
                if (elect_one()) load_resident_weights(s + 1);
                __syncwarp();
            }
        }
    } else if (MODE == MODE_GATHER && warp < 4) {
        // ============================================================ A producers (cp.async gather)
        // One stage = `tap_group` filter taps x 16 input channels (2 planes) x 128 rows: each of the
        // 128 producer threads owns one tile row and issues 2 x tap_group 16-byte cp.async per stage
        // (zero-fill outside the image), so up to (kGatherLag+1) x 18 loads are in flight per thread.
        pdl_wait();
        const __half* in0 = p.seg[0].in;
        const __half* in0_lo = p.seg[0].in_lo;
        const int r = threadIdx.x;                 // row of the tile
        const int Wp_o = p.Wout + 2, HpWp_o = (p.Hout + 2) * Wp_o;
        const int Wp_i = p.Win + 2, HpWp_i = (p.Hin + 2) * Wp_i;
        const int TG = p.tap_group, n_groups = taps / TG;
        uint32_t it = 0, g_slot = 0, g_ph = 0, g_lag = 0;      // g_lag: slot of the stage kGatherLag iterations back
        for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
            const long long q = (long long)tile * kTileM + r;
            const int n = int(q / HpWp_o);
            const int rem = int(q - (long long)n * HpWp_o);
            const int yo = rem / Wp_o - 1, xo = rem % Wp_o - 1;
            const bool interior = q < p.P_out && yo >= 0 && yo < p.Hout && xo >= 0 && xo < p.Wout;
            const long long crop0 = (long long)n * HpWp_i;
            for (int tg = 0; tg < n_groups; ++tg) {
                // source pixel of every tap of this group, computed once and reused by all K chunks
                long long pix[9];
                uint32_t okmask = 0;
#pragma unroll
                for (int tt = 0; tt < 9; ++tt) {
                    const int t = tg * TG + tt;
                    const int yi = yo * p.stride + t / p.kw - p.pad;
                    const int xi = xo * p.stride + t % p.kw - p.pad;
                    const bool ok = tt < TG && interior && yi >= 0 && yi < p.Hin && xi >= 0 && xi < p.Win;
                    pix[tt] = ok ? (crop0 + (yi + 1) * Wp_i + (xi + 1)) * 8 : 0;
                    okmask |= ok ? (1u << tt) : 0u;
                }
                for (int c = 0; c < p.n_chunks; ++c, ++it) {
                    const uint32_t slot = g_slot, ph = g_ph;
                    if (++g_slot == uint32_t(p.stages)) { g_slot = 0; g_ph ^= 1u; }
                    MBAR_WAIT(smem_u32(&hdr->empty[slot]), ph ^ 1u);
                    const uint32_t st = smem_u32(stage0 + size_t(slot) * stage_bytes) + uint32_t(r) * 16u;
                    const __half* src0 = in0 + (long long)(c * 2) * p.in_plane * 8;
#pragma unroll
                    for (int tt = 0; tt < 9; ++tt) {
                        if (tt < TG) {
                            const bool ok = (okmask >> tt) & 1u;
                            const __half* src = src0 + pix[tt];
                            cp_async16(st + uint32_t(tt) * 4096u, src, ok);
                            cp_async16(st + uint32_t(tt) * 4096u + 2048u, src + p.in_plane * 8, ok);
                        }
                    }
                    if (SPLIT) {            // the lo planes of all taps of the group land behind the hi block
                        const __half* srl0 = in0_lo + (long long)(c * 2) * p.in_plane * 8;
                        const uint32_t stl = st + uint32_t(TG) * 4096u;
#pragma unroll
                        for (int tt = 0; tt < 9; ++tt) {
                            if (tt < TG) {
                                const bool ok = (okmask >> tt) & 1u;
                                const __half* src = srl0 + pix[tt];
                                cp_async16(stl + uint32_t(tt) * 4096u, src, ok);
                                cp_async16(stl + uint32_t(tt) * 4096u + 2048u, src + p.in_plane * 8, ok);
                            }
                        }
                    }
                    cp_async_commit();
                    if (it >= uint32_t(kGatherLag)) {
                        cp_async_wait<kGatherLag>();
                        fence_proxy_async_smem();
                        mbar_arrive(smem_u32(&hdr->full[g_lag]));
                        if (++g_lag == uint32_t(p.stages)) g_lag = 0;
                    }
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async_smem();
        for (uint32_t d = (it > uint32_t(kGatherLag) ? it - kGatherLag : 0u); d < it; ++d) {
            mbar_arrive(smem_u32(&hdr->full[g_lag]));
            if (++g_lag == uint32_t(p.stages)) g_lag = 0;
        }
    } else if (MODE == MODE_GATHER && warp == R::kWWarp) {
        // ============================================================ W producer (bulk copies)
        if (p.w_resident) {
            if (elect_one()) load_resident_weights(0);
            __syncwarp();
        } else {
            const __half* wg = p.seg[0].w + size_t(nb) * p.n_tile * 8;
            uint32_t w_slot = 0, w_ph = 0;
            const int TG = p.tap_group, n_groups = taps / TG;
            for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x)
                for (int tg = 0; tg < n_groups; ++tg)
                    for (int c = 0; c < p.n_chunks; ++c) {
                        const uint32_t slot = w_slot, ph = w_ph;
                        if (++w_slot == uint32_t(p.stages)) { w_slot = 0; w_ph ^= 1u; }
                        MBAR_WAIT(smem_u32(&hdr->empty[slot]), ph ^ 1u);
                        if (elect_one()) {
                            const uint32_t bar = smem_u32(&hdr->full[slot]);
                            mbar_arrive_expect_tx(bar, uint32_t(p.w_stage_bytes));
                            const uint32_t ws = smem_u32(stage0 + size_t(slot) * stage_bytes) + p.a_stage_bytes;
                            load_stage_weights(wg, ws, bar, tg * TG, tg * TG + TG, c);
                            if (SPLIT) load_stage_weights(wg + w_lo_off, ws + (p.w_stage_bytes >> 1), bar, tg * TG, tg * TG + TG, c);
                        }
                        __syncwarp();
                    }
        }
    } else if (warp == R::kMmaWarp || warp == R::kMmaWarp + 1) {
        // ============================================================ MMA issuers (warp-convergent, one elected lane issues)
        // Two warps take alternate tiles, one stage ring each, so one warp's barrier round trips overlap
        // the other's MMAs (profiles/r01_conv_role_bench.csv).
        const uint32_t mw = uint32_t(warp - R::kMmaWarp);
        const uint32_t idesc = umma_idesc_f16(kTileM, uint32_t(p.n_tile) * (NCAT ? 2u : 1u));
        const uint32_t idesc_half = umma_idesc_f16(kTileM, uint32_t(p.n_tile));      // N concatenation: the x_lo W_hi product
        // descriptor = hi:lo, lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version<<14: only lo changes
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lbo = (uint32_t(p.a_plane_bytes) >> 4) << 16, b_lbo = (slab_bytes >> 4) << 16;
        const uint32_t a_kstep = (2u * uint32_t(p.a_plane_bytes)) >> 4, b_kstep = (2u * slab_bytes) >> 4;
        const int ksteps = planes_per_chunk >> 1;
        const int Wp = p.Wout + 2;
        // descriptor-unit (16 B) pitch between the weight slabs of consecutive taps
        const uint32_t w_tap_stride = MODE == MODE_LINEAR
            ? ((p.w_resident ? uint32_t(cin8) : uint32_t(planes_per_chunk)) * slab_bytes) >> 4
            : ((p.w_resident ? uint32_t(cin8) : 2u) * slab_bytes) >> 4;
        const bool active = p.rings == 2 || mw == 0u;       // a single ring is served by warp 0
        const uint32_t tile_bytes = uint32_t(p.tile_stride) * 16u;
        // where filter tap t starts inside the landed run, in 16-byte descriptor units.  Plain 3x3: (r-1) rows + (s-1)
        // pixels.  Phase-split input of a stride-2 conv (in_s2d): input row 2y + r - 1 is row y - 1 of the odd-row phase
        // for r = 0, row y of the even phase for r = 1, row y of the odd phase for r = 2 (columns alike), so tap (r, s)
        // reads phase block (r != 1) * 2 + (s != 1) at a shift of -1 / 0 rows and pixels: the nine taps are nine
        // (phase block, shift) pairs over the SAME halo run, and the weights keep their [tap][Cin/8][Cout][8] layout.
        const uint32_t phase_units = (uint32_t(planes_per_chunk) * uint32_t(p.a_plane_bytes)) >> 4;
        uint32_t sh[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            int v = 0;
            if (MODE == MODE_LINEAR) {
                const int r = t / 3, s_ = t % 3;
                if (p.in_s2d) v = (r == 0 ? -Wp : 0) + (s_ == 0 ? -1 : 0) + int(((r != 1) * 2 + (s_ != 1)) * phase_units);
                else if (taps == 9) v = (r - 1) * Wp + (s_ - 1);
                else if (taps == 3) v = (t - 1) * Wp;
            }
            sh[t] = uint32_t(v);
        }
        uint32_t tl = 0, ul = 0;
        uint32_t m_slot = 0, m_ph = 0;                  // this issuer's position in its stage ring
        uint32_t a_buf = 0, a_par = 1u;                 // accumulator ring position of the next tile (all tiles, both issuers)
        const bool prof = (p.debug & 32) != 0;
        long long pt_acc = 0, pt_full = 0, pt_issue = 0, pt_units = 0, pt_t0 = prof ? clock64() : 0, pt_mark = 0;
        for (int s = 0; s < n_segs; ++s) {
            const uint32_t w_res_u32 = smem_u32(w_res) + uint32_t(wbuf_of(s)) * uint32_t(p.w_res_bytes);
            if (p.w_resident && active) MBAR_WAIT(smem_u32(&hdr->w_ready[wbuf_of(s)]), uint32_t(wuse_of(s)) & 1u);
            for (int ju = 0; ju < my_units; ++ju) {
                const int gcount = unit_tiles(unit_of(ju));
                tl += uint32_t(gcount);
                const uint32_t ul0 = ul++;
                const uint32_t b0 = a_buf, par0 = a_par;      // first accumulator of this unit; advance the ring past the unit
                a_buf += uint32_t(gcount);
                if (a_buf >= nacc) { a_buf -= nacc; a_par ^= 1u; }
                if (p.rings == 2 ? (ul0 & 1u) != mw : mw != 0u) continue;   // one ring per issuer
                // accumulator of every tile of the unit (all must have been drained); computed here, outside the
                // elected region, so that the MMA descriptors below stay pure uniform-register arithmetic
                uint32_t dt[4], bufg[4];
                if (prof) { pt_mark = clock64(); ++pt_units; }
                {
                    uint32_t b = b0, par = par0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        bufg[g] = b;
                        dt[g] = tmem_base + b * buf_cols;
                        if (g < gcount) MBAR_WAIT(smem_u32(&hdr->tmem_empty[b]), par);
                        if (++b == nacc) { b = 0; par ^= 1u; }
                    }
                }
                tc_fence_after();
                if (prof) { const long long t = clock64(); pt_acc += t - pt_mark; pt_mark = t; }
                for (int ki = 0; ki < kiters; ++ki) {
                    const uint32_t slot = (p.rings == 2 ? mw * p.stages : 0u) + m_slot, ph = m_ph;
                    if (++m_slot == uint32_t(p.stages)) { m_slot = 0; m_ph ^= 1u; }
                    MBAR_WAIT(smem_u32(&hdr->full[slot]), ph);
                    tc_fence_after();
                    if (prof) { const long long t = clock64(); pt_full += t - pt_mark; pt_mark = t; }
                    if (tl_on && mw == 0u && ul0 == 0u && ki == 0 && lane == 0) hdr->stamps[2] = global_timer_ns();
                    if (elect_one()) {
                        const uint32_t a_base0 = smem_u32(stage0) + slot * uint32_t(stage_bytes);
                        const uint32_t w_stage = a_base0 + p.a_stage_bytes;
                        const uint32_t acc = ki > 0 ? 1u : 0u;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                        if (g >= gcount) break;
                        const uint32_t d_tmem = dt[g];
                        const uint32_t a_base = a_base0 + uint32_t(g) * tile_bytes;     // tile g of the run
                        if (MODE == MODE_LINEAR) {
                            const uint32_t w0 = p.w_resident ? w_res_u32 + uint32_t(ki * planes_per_chunk) * slab_bytes : w_stage;
                            const uint32_t a_lo = ((a_base + uint32_t(p.halo) * 16u) >> 4) | a_lbo;
                            const uint32_t b_lo = (w0 >> 4) | b_lbo;
                            if (!(p.debug & 2)) {
#define POCO_ISSUE(T, K)                                                                                                    \
    do {                                                                                                                    \
        if (NCAT) {         /* x_hi [W_hi | W_lo] into both column groups, then x_lo W_hi onto the first */                 \
            issue_linear<T, K>(d_tmem, a_lo, b_lo, sh, a_kstep, b_kstep, w_tap_stride, desc_hi, idesc, acc);                \
            issue_linear<T, K>(d_tmem, a_lo2, b_lo, sh, a_kstep, b_kstep, w_tap_stride, desc_hi, idesc_half, 1u);           \
            break;                                                                                                          \
        }                                                                                                                   \
        if (SPLIT) {        /* x_lo W_hi + x_hi W_lo first: the tensor core truncates every accumulation to the magnitude */ \
                            /* of the running sum, so the 2^-11-sized correction terms go in while it is still small     */ \
            issue_linear<T, K>(d_tmem, a_lo2, b_lo, sh, a_kstep, b_kstep, w_tap_stride, desc_hi, idesc, acc);               \
            issue_linear<T, K>(d_tmem, a_lo, b_lo2, sh, a_kstep, b_kstep, w_tap_stride, desc_hi, idesc, 1u);                \
        }                                                                                                                   \
        issue_linear<T, K>(d_tmem, a_lo, b_lo, sh, a_kstep, b_kstep, w_tap_stride, desc_hi, idesc, SPLIT ? 1u : acc);       \
    } while (0)
                                const uint32_t a_lo2 = a_lo + (p.in_s2d ? 4u : 1u) * phase_units;
                                const uint32_t b_lo2 = b_lo + ((p.w_resident ? uint32_t(p.w_res_bytes) : uint32_t(p.w_stage_bytes)) >> 5);
                                if (taps == 9) {
                                    switch (ksteps) {
                                        case 1: POCO_ISSUE(9, 1); break;
                                        case 2: POCO_ISSUE(9, 2); break;
                                        case 3: POCO_ISSUE(9, 3); break;
                                        default: POCO_ISSUE(9, 4); break;
                                    }
                                } else if (DXN) {
                                    switch (ksteps) {
                                        case 1: POCO_ISSUE(3, 1); break;
                                        case 2: POCO_ISSUE(3, 2); break;
                                        case 3: POCO_ISSUE(3, 3); break;
                                        default: POCO_ISSUE(3, 4); break;
                                    }
                                } else {
                                    switch (ksteps) {
                                        case 1: POCO_ISSUE(1, 1); break;
                                        case 2: POCO_ISSUE(1, 2); break;
                                        case 3: POCO_ISSUE(1, 3); break;
                                        default: POCO_ISSUE(1, 4); break;
                                    }
                                }
#undef POCO_ISSUE
                            }
                        } else {
                            const int tg = ki / p.n_chunks, c = ki - tg * p.n_chunks, TG = p.tap_group;
                            const uint32_t w0 = p.w_resident ? w_res_u32 + uint32_t(tg * TG * cin8 + c * 2) * slab_bytes : w_stage;
                            const uint32_t a_lo = (a_base >> 4) | a_lbo;
                            const uint32_t b_lo = (w0 >> 4) | b_lbo;
                            const uint32_t a_lo2 = a_lo + uint32_t(TG) * 256u;       // lo block: TG taps x 4 KB behind the hi block
                            const uint32_t b_lo2 = b_lo + ((p.w_resident ? uint32_t(p.w_res_bytes) : uint32_t(p.w_stage_bytes)) >> 5);
#define POCO_ISSUE_G(T)                                                                     \
    do {                                                                                    \
        if (SPLIT) {                                                                        \
            issue_gather<T>(d_tmem, a_lo2, b_lo, w_tap_stride, desc_hi, idesc, acc);        \
            issue_gather<T>(d_tmem, a_lo, b_lo2, w_tap_stride, desc_hi, idesc, 1u);         \
        }                                                                                   \
        issue_gather<T>(d_tmem, a_lo, b_lo, w_tap_stride, desc_hi, idesc, SPLIT ? 1u : acc); \
    } while (0)
                            switch (TG) {
                                case 9: POCO_ISSUE_G(9); break;
                                case 7: POCO_ISSUE_G(7); break;
                                case 3: POCO_ISSUE_G(3); break;
                                default: POCO_ISSUE_G(1); break;
                            }
#undef POCO_ISSUE_G
                        }
                        }
                        umma_commit(smem_u32(&hdr->empty[slot]));      // smem slot free once these MMAs retire
                        if (ki == kiters - 1)                             // accumulators complete
                        {
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                if (g < gcount) umma_commit(smem_u32(&hdr->tmem_full[bufg[g]]));
                        }
                    }
                    __syncwarp();
                    if (prof) { const long long t = clock64(); pt_issue += t - pt_mark; pt_mark = t; }
                    if (tl_on && mw == 0u && ul0 == 0u && ki == kiters - 1 && lane == 0) hdr->stamps[3] = global_timer_ns();
                }
            }
            if (p.w_resident && active && n_segs > 1) {     // this warp's MMAs of the segment no longer read the weight buffer
                if (elect_one()) umma_commit(smem_u32(&hdr->w_free[wbuf_of(s)]));
                __syncwarp();
            }
        }
        if (prof && lane == 0 && p.prof != nullptr) {
            unsigned long long* o = p.prof + mw * 8;
            atomicAdd(o + 0, (unsigned long long)(clock64() - pt_t0));
            atomicAdd(o + 1, (unsigned long long)pt_acc);
            atomicAdd(o + 2, (unsigned long long)pt_full);
            atomicAdd(o + 3, (unsigned long long)pt_issue);
            atomicAdd(o + 4, (unsigned long long)pt_units);
            atomicAdd(o + 5, 1ull);
        }
    } else if (warp >= R::kEpiWarp0) {
        // ============================================================ epilogue (8 autonomous warps)
        // Work item = IPL output planes (8 accumulator columns each) of one tile; the two halves of
        // the epilogue take alternate items.  Inside a half each of the 4 warps owns the 32 tile rows
        // of its TMEM lane group and runs on its own: no CTA-level barrier.  Per warp: the residual
        // slice (32 rows x 16 B per plane = one contiguous 512-byte run of the planar layout) arrives
        // by bulk copy into a private ring, prefetched `res_ring` items ahead; results are written
        // straight from registers, one 16-byte store per lane and plane = 512 contiguous bytes per
        // warp instruction (a shared-memory staged bulk-store variant measured slower: its
        // fence.proxy.async + store-queue waits sat on the per-item critical path).  Halo rows are
        // never written (zero-halo invariant).
        pdl_wait();
        const int ew = warp - R::kEpiWarp0;
        const int half = ew >> 2;
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int row = lg * 32 + lane;
        const int Wp_o = p.Wout + 2, HpWp_o = (p.Hout + 2) * Wp_o;
        const int n_out = p.n_out;
        const int plane0 = DXN ? 0 : (nb * p.n_tile) >> 3;
        const int items = p.epi_items;                  // work items per tile
        // (tile, item) pairs are dealt alternately to the two halves: with one item per tile (N <= 32) the
        // halves take alternate TILES, so each warp has two tile-times for its waits and index math
        const int odd = items & 1;
        auto k_first = [&](uint32_t tl_) { return (half + (odd ? int(tl_ & 1u) : 0)) & 1; };
        const bool skip_store = (p.debug & 8) != 0;
        constexpr uint32_t kSlot = IPL * 512 * SP;      // one item slice of this warp: IPL planes x 32 rows x 16 B (split: hi planes, then lo planes)
        uint8_t* res_ring = smem + kHeaderBytes + ew * warp_ring_bytes;
        unsigned long long* res_full = hdr->res_full + ew * kMaxResRing;
        const uint32_t rr_n = uint32_t(res_ring_n > 0 ? res_ring_n : 1);      // (0 = no ring: a conv without residual)
        // Residual prefetch cursor (used by the elected lane): the next residual-bearing (segment, tile,
        // item) of this warp.  A segment's residual may be the output of the segment two before it, so
        // the cursor never runs past segment `cur + 1` (everything up to `cur - 1` is complete on this
        // CTA: all epilogue warps meet at a barrier between segments).
        auto seg_has_res = [&](int s_) { return p.seg[s_].res != nullptr && !(p.debug & 16); };
        int pf_s = 0;
        while (pf_s < n_segs && !seg_has_res(pf_s)) ++pf_s;
        int pf_j = 0, pf_k = k_first(uint32_t(pf_s * my_tiles));
        uint32_t pf_tl = uint32_t(pf_s * my_tiles);
        const __half* pf_res = pf_s < n_segs ? p.seg[pf_s].res : nullptr;
        const __half* pf_res_lo = SPLIT ? p.seg[0].res_lo : nullptr;
        uint32_t pf_issued = 0, pf_slot = 0;
        auto prefetch_residual = [&](int cur_seg, uint32_t upto) {      // elected lane: top the ring up to `upto` items
            while (pf_issued < upto) {
                while (pf_k >= items) {         // advance to the next tile (segment) with an item for this half
                    ++pf_j;
                    ++pf_tl;
                    if (pf_j >= my_tiles) {
                        do { ++pf_s; } while (pf_s < n_segs && !seg_has_res(pf_s));
                        if (pf_s >= n_segs) { pf_k = 0; pf_s = n_segs; return; }
                        pf_j = 0;
                        pf_tl = uint32_t(pf_s * my_tiles);
                        pf_res = p.seg[pf_s].res;
                    }
                    pf_k = k_first(pf_tl);
                }
                if (pf_s >= n_segs || pf_s > cur_seg + 1) return;
                const long long tile_ = tile_of(pf_j);
                const int item = pf_k;
                pf_k += 2;
                const uint32_t slot = pf_slot;
                if (++pf_slot == rr_n) pf_slot = 0;
                ++pf_issued;
                const long long qw = tile_ * p.tile_stride + p.tile_origin + lg * 32;
                const long long left = p.P_out - qw;
                const uint32_t bar = smem_u32(&res_full[slot]);
                if (left <= 0) {                // rows past the end of the tensor: complete the slot's phase anyway
                    mbar_arrive(bar);
                    continue;
                }
                const uint32_t rows = uint32_t(left < 32 ? left : 32);
                const int planes = min(ipl, (n_out >> 3) - item * ipl);
                mbar_arrive_expect_tx(bar, uint32_t(planes * SP) * rows * 16u);
                const long long res_off = ((long long)(plane0 + item * ipl) * p.res_plane + qw) * 8;
                const __half* src = pf_res + res_off;
                const uint32_t dst = smem_u32(res_ring) + slot * kSlot;
                for (int pl = 0; pl < planes; ++pl, src += p.res_plane * 8)
                    bulk_g2s(dst + uint32_t(pl) * 512u, src, rows * 16u, bar);
                if (SPLIT) {
                    const __half* srl = pf_res_lo + res_off;
                    for (int pl = 0; pl < planes; ++pl, srl += p.res_plane * 8)
                        bulk_g2s(dst + uint32_t(IPL + pl) * 512u, srl, rows * 16u, bar);
                }
            }
        };
        // position of a row inside its crop, advanced incrementally: tile -> next tile of the unit, and last
        // tile of a unit -> first tile of this CTA's next unit (only the globally last unit can be short)
        const uint32_t step_in = uint32_t(p.tile_stride % HpWp_o);
        const uint32_t step_unit = uint32_t((((long long)gridDim.x * G - (G - 1)) * p.tile_stride) % HpWp_o);
        const uint32_t magic_w = 0xFFFFFFFFu / uint32_t(Wp_o) + 1u;      // exact floor(n / Wp) for n < 2^16
        const uint32_t magic_hw = 0xFFFFFFFFu / uint32_t(HpWp_o) + 1u;   // exact n / HpWp for exact multiples n < 2^32
        uint32_t tl = 0, g = 0;                         // g counts residual items consumed
        uint32_t e_buf = 0, e_par = 0;                  // accumulator ring position of tile `tl`
        uint32_t rs_slot = 0, rs_par = 0;               // residual ring position of item `g`
        int xbuf = 0;                                   // dx-in-N exchange buffer of this half (alternates per item)
        for (int s = 0; s < n_segs; ++s) {
            const ChainSeg& sg = p.seg[s];
            const bool has_res = seg_has_res(s);
            const int relu = sg.relu;
            int* flags_cur = s + 1 < n_segs ? p.flags + size_t(s) * p.num_m_tiles : nullptr;
            float* bias_s = hdr->bias[s & 1];
            {   // this segment's bias -> shared memory; between segments every epilogue warp has finished
                // (and fenced, in its last signal_tiles) the stores of the previous one, which also bounds the
                // residual prefetch
                const int t = threadIdx.x - R::kEpiWarp0 * 32;
                for (int i = t; i < n_out; i += 256) bias_s[i] = sg.bias[(DXN ? 0 : nb * p.n_tile) + i];
                named_barrier_sync(1, 256);
            }
            if (elect_one()) {
                fence_proxy_async_all();
                prefetch_residual(s, g + rr_n);
            }
            __syncwarp();
            // position of this thread's row inside its crop, advanced incrementally from tile to tile
            uint32_t rem = uint32_t((((long long)blockIdx.x << g_log2) * p.tile_stride + p.tile_origin + row + HpWp_o) % HpWp_o);
            // Completion flags are published in batches: one gpu-scope fence (it waits for the warp's
            // outstanding stores, ~1 us) covers the last kSignalEvery tiles this warp had a share of.
            // Consumers run a whole segment behind, so the delay costs nothing; the segment end flushes.
            const int kSignalEvery = max(1, min(8, my_tiles / (odd ? 8 : 4)));
            int sig_first = 0, sig_owned = 0;               // local tiles [sig_first, j] still unpublished
            const uint32_t tl_seg0 = tl;
            auto signal_tiles = [&](int j_last) {
                if (!(p.debug & 128)) fence_acq_rel_gpu();          // (128: timing experiment only -- no ordering)
                __syncwarp();
                const int jl = sig_first + lane;
                if (jl <= j_last && k_first(tl_seg0 + uint32_t(jl)) < items)
                    red_relaxed_gpu_add(flags_cur + tile_of(jl), 1);
                sig_first = j_last + 1;
                sig_owned = 0;
            };
            int j = -1;
            for (int ju = 0; ju < my_units; ++ju)
            for (int g_ = 0, tile = unit_of(ju) * G; g_ < unit_tiles(unit_of(ju)); ++g_, ++tile, ++tl) {
                ++j;
                const uint32_t buf = e_buf, buf_par = e_par;
                if (++e_buf == nacc) { e_buf = 0; e_par ^= 1u; }
                const long long qw = (long long)tile * p.tile_stride + p.tile_origin + lg * 32;     // first row of this warp's slice
                const uint32_t yy = __umulhi(rem, magic_w), xx = rem - yy * uint32_t(Wp_o);
                const bool interior = qw + lane < p.P_out && yy >= 1u && yy <= uint32_t(p.Hout) && xx >= 1u && xx <= uint32_t(p.Wout) &&
                                      !(DXN && (row == 0 || row == kTileM - 1));     // (dx-in-N: the tile's edge rows belong to its neighbours)
                long long s2d_off = 0;      // this row's pixel in the phase-split second output (plane 0 of its phase)
                if (S2DOUT && interior) {
                    const uint32_t n_crop = __umulhi(uint32_t(qw + lane) - rem, magic_hw);      // (q - rem) = crop * HpWp
                    const uint32_t y = yy - 1u, x = xx - 1u;
                    const uint32_t ph = (y & 1u) * 2u + (x & 1u);
                    s2d_off = ((long long)(ph * uint32_t(p.Cout >> 3)) * p.s2d_plane + (long long)n_crop * p.s2d_HpWp +
                               ((y >> 1) + 1u) * uint32_t(p.s2d_Wp) + (x >> 1) + 1u) * 8;
                }
                rem += (g_ + 1 == G) ? step_unit : step_in;
                if (rem >= uint32_t(HpWp_o)) rem -= uint32_t(HpWp_o);
                const long long left = p.P_out - qw;
                const uint32_t rows_w = left <= 0 ? 0u : uint32_t(left < 32 ? left : 32);
                const int k0 = k_first(tl);
                if (k0 >= items) continue;                  // the other half drains this tile
                MBAR_WAIT(smem_u32(&hdr->tmem_full[buf]), buf_par);
                tc_fence_after();
                if (tl_on && ew == 0 && j == 0 && lane == 0) hdr->stamps[4] = global_timer_ns();
                if (p.debug & 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&hdr->tmem_empty[buf]));
                    if (flags_cur != nullptr && (++sig_owned >= kSignalEvery)) signal_tiles(j);
                    continue;
                }
                const uint32_t taddr = tmem_base + buf * buf_cols + (uint32_t(lg * 32) << 16);
                for (int item = k0; item < items; item += 2) {
                    const int c0 = item * ipl * 8;
                    const int planes = min(ipl, (n_out - c0) >> 3);         // 2 or 4
                    uint32_t v[IPL * 8];
                    if (DXN && IPL == 4) {
                        // the three column groups D_0 | D_1 | D_2 of these 32 output channels
                        uint32_t d0[32], d1[32], d2[32];
                        tmem_ld16(taddr + uint32_t(c0), d0);
                        tmem_ld16(taddr + uint32_t(c0 + 16), d0 + 16);
                        tmem_ld16(taddr + uint32_t(n_out + c0), d1);
                        tmem_ld16(taddr + uint32_t(n_out + c0 + 16), d1 + 16);
                        tmem_ld16(taddr + uint32_t(2 * n_out + c0), d2);
                        tmem_ld16(taddr + uint32_t(2 * n_out + c0 + 16), d2 + 16);
                        tmem_ld_wait();
                        if (item + 2 >= items) {            // this warp is done with the accumulator buffer
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&hdr->tmem_empty[buf]));
                        }
                        // rows 31 / 0 of this warp are the left / right neighbours of the adjacent warps' edge rows
                        float(*xs)[2][32] = hdr->xch[half][xbuf];
                        xbuf ^= 1;
                        if (lane == 31) {
#pragma unroll
                            for (int c = 0; c < 32; c += 4)
                                *reinterpret_cast<uint4*>(&xs[lg][0][c]) = make_uint4(d0[c], d0[c + 1], d0[c + 2], d0[c + 3]);
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int c = 0; c < 32; c += 4)
                                *reinterpret_cast<uint4*>(&xs[lg][1][c]) = make_uint4(d2[c], d2[c + 1], d2[c + 2], d2[c + 3]);
                        }
                        named_barrier_sync(2 + half, 128);
                        // every lane reads the neighbour warps' edge rows (broadcast loads, no divergence) and
                        // selects them only at its own edge lanes; rows 0 / 127 of the tile are never stored
                        const float* nl = xs[lg > 0 ? lg - 1 : 0][0];
                        const float* nr = xs[lg < 3 ? lg + 1 : 3][1];
#pragma unroll
                        for (int c4 = 0; c4 < 32; c4 += 4) {
                            const float4 l4 = *reinterpret_cast<const float4*>(nl + c4);
                            const float4 r4 = *reinterpret_cast<const float4*>(nr + c4);
                            const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int c = c4 + i;
                                float a = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[c]), 1);
                                float b = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[c]), 1);
                                a = lane == 0 ? lv[i] : a;
                                b = lane == 31 ? rv[i] : b;
                                v[c] = __float_as_uint(a + __uint_as_float(d1[c]) + b);
                            }
                        }
                    } else {
                        tmem_ld16(taddr + uint32_t(c0), v);
                        if (IPL == 4 && planes == 4) tmem_ld16(taddr + uint32_t(c0 + 16), v + (IPL == 4 ? 16 : 0));
                        if (NCAT) {             // second column group: the x_hi W_lo products of the same output channels
                            uint32_t v2[IPL * 8];
                            tmem_ld16(taddr + uint32_t(n_out + c0), v2);
                            if (IPL == 4 && planes == 4) tmem_ld16(taddr + uint32_t(n_out + c0 + 16), v2 + (IPL == 4 ? 16 : 0));
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < IPL * 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
                        } else {
                            tmem_ld_wait();
                        }
                        if (item + 2 >= items) {            // this warp is done with the accumulator buffer
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&hdr->tmem_empty[buf]));
                        }
                    }
                    const uint32_t rslot = rs_slot;
                    if (has_res) MBAR_WAIT(smem_u32(&res_full[rslot]), rs_par);
                    const long long out_off = ((long long)(plane0 + item * ipl) * p.out_plane + qw + lane) * 8;
                    __half* outp = sg.out + out_off;
                    const uint8_t* rb = res_ring + rslot * kSlot + lane * 16;
#pragma unroll
                    for (int pl = 0; pl < IPL; ++pl) {
                        if (IPL == 4 && pl >= planes) break;
                        const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[c0 + pl * 8]);
                        const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[c0 + pl * 8 + 4]);
                        float f[8] = {__uint_as_float(v[pl * 8 + 0]) + b0.x, __uint_as_float(v[pl * 8 + 1]) + b0.y,
                                      __uint_as_float(v[pl * 8 + 2]) + b0.z, __uint_as_float(v[pl * 8 + 3]) + b0.w,
                                      __uint_as_float(v[pl * 8 + 4]) + b1.x, __uint_as_float(v[pl * 8 + 5]) + b1.y,
                                      __uint_as_float(v[pl * 8 + 6]) + b1.z, __uint_as_float(v[pl * 8 + 7]) + b1.w};
                        if (relu == 2) {                    // ReLU before the residual add (hrnet_cls.py:473-474)
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
                        }
                        if (has_res) {
                            const uint4 r4 = *reinterpret_cast<const uint4*>(rb + pl * 512);
                            const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 t2 = unpack_half2(rr[i]);
                                f[2 * i] += t2.x;
                                f[2 * i + 1] += t2.y;
                            }
                            if (SPLIT) {
                                const uint4 l4 = *reinterpret_cast<const uint4*>(rb + (IPL + pl) * 512);
                                const uint32_t rl[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 t2 = unpack_half2(rl[i]);
                                    f[2 * i] += t2.x;
                                    f[2 * i + 1] += t2.y;
                                }
                            }
                        }
                        if (relu == 1) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
                        }
                        if (interior && !skip_store) {      // 32 lanes x 16 B = one contiguous 512-byte run of the plane
                            uint4 o4;
                            o4.x = pack_half2(f[0], f[1]); o4.y = pack_half2(f[2], f[3]);
                            o4.z = pack_half2(f[4], f[5]); o4.w = pack_half2(f[6], f[7]);
                            if (!S2DOUT || !p.s2d_only) *reinterpret_cast<uint4*>(outp + (long long)pl * p.out_plane * 8) = o4;
                            uint4 l4 = make_uint4(0, 0, 0, 0);
                            if (SPLIT) {        // lo = fp16(y - hi)
                                const uint32_t oh[4] = {o4.x, o4.y, o4.z, o4.w};
                                uint32_t ol[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 h2 = unpack_half2(oh[i]);
                                    ol[i] = pack_half2(f[2 * i] - h2.x, f[2 * i + 1] - h2.y);
                                }
                                l4 = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                                if (!S2DOUT || !p.s2d_only) *reinterpret_cast<uint4*>(sg.out_lo + out_off + (long long)pl * p.out_plane * 8) = l4;
                            }
                            if (S2DOUT) {       // the same pixel in the phase-split copy (feeds a stride-2 conv)
                                const long long so = s2d_off + (long long)(plane0 + item * ipl + pl) * p.s2d_plane * 8;
                                *reinterpret_cast<uint4*>(p.s2d_out + so) = o4;
                                if (SPLIT) *reinterpret_cast<uint4*>(p.s2d_out_lo + so) = l4;
                            }
                        }
                    }
                    if (has_res) {
                        ++g;
                        if (++rs_slot == rr_n) { rs_slot = 0; rs_par ^= 1u; }
                        __syncwarp();                       // every lane is done with the residual slot
                        if (elect_one()) prefetch_residual(s, g + rr_n);
                        __syncwarp();
                    }
                }
                if (tl_on && ew == 0 && lane == 0) { if (j == 0) hdr->stamps[5] = global_timer_ns(); hdr->stamps[6] = global_timer_ns(); }
                if (flags_cur != nullptr && (++sig_owned >= kSignalEvery)) signal_tiles(j);
            }
            if (flags_cur != nullptr) signal_tiles(my_tiles - 1);
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == R::kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
    }
    if (tl_on && threadIdx.x == 0 && p.prof != nullptr) {
        hdr->stamps[7] = global_timer_ns();
        const unsigned long long slot = atomicAdd(p.prof + 127, 1ull);
        if (slot < 12)
            for (int i = 0; i < 8; ++i) p.prof[16 + slot * 8 + i] = hdr->stamps[i];
    }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

int64_t conv_flops(const poco_conv* d) {
    return 2ll * d->out.N * d->out.H * d->out.W * d->out.C * (d->in_s2d ? d->in.C / 4 : d->in.C) * d->kh * d->kw;
}

int conv_tc_launch(const poco_conv* d, cudaStream_t s) { return conv_tc_launch_chain(d, 1, nullptr, s); }

// segs[0..n): same geometry (3x3/s1/p1 or 1x1/s1 when n > 1), flags: (n-1) * num_m_tiles zeroed ints
int conv_tc_launch_chain(const poco_conv* segs, int n_segs, int32_t* flags, cudaStream_t s) {
    const poco_conv* d = segs;
    const poco_act &in = d->in, &out = d->out;
    POCO_CHECK(n_segs >= 1 && n_segs <= kMaxSegs, "bad chain length");
    POCO_CHECK(in.C % 16 == 0 && out.C % 16 == 0, "Cin and Cout must be multiples of 16");
    POCO_CHECK(in.N == out.N, "batch mismatch");
    POCO_CHECK(d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
    const bool in_s2d = d->in_s2d != 0;
    if (in_s2d) {
        POCO_CHECK(n_segs == 1 && d->kh == 3 && d->kw == 3 && d->stride == 2 && d->pad == 1 && d->wfmt == 0 && in.C % 64 == 0 &&
                       in.H == out.H && in.W == out.W,
                   "in_s2d: a 3x3 / stride 2 / pad 1 conv whose input is given phase-split (4 Cin channels at the output resolution)");
    } else {
        POCO_CHECK((in.H + 2 * d->pad - d->kh) / d->stride + 1 == out.H && (in.W + 2 * d->pad - d->kw) / d->stride + 1 == out.W,
                   "output geometry does not match the convolution");
    }
    if (d->out_s2d.data != nullptr) {
        const poco_act& s = d->out_s2d;
        POCO_CHECK(n_segs == 1 && d->wfmt != 1 && out.H % 2 == 0 && out.W % 2 == 0 && s.C == 4 * out.C && s.N == out.N && s.H == out.H / 2 &&
                       s.W == out.W / 2 && (s.lo != nullptr) == (in.lo != nullptr),
                   "out_s2d: 4 Cout channels at half the (even) output resolution, same precision mode");
    }
    POCO_CHECK(!d->s2d_only || d->out_s2d.data != nullptr, "s2d_only without out_s2d");
    POCO_CHECK((kTileM + out.W + 3) * 16 <= POCO_ACT_GUARD_BYTES, "tile halo exceeds the activation guard");

    const bool split = in.lo != nullptr;          // split-precision ("parity") mode: see the SPLIT template parameter
    const int sp = split ? 2 : 1;
    POCO_CHECK(!split || ((out.lo != nullptr || d->s2d_only) && n_segs == 1 && (d->wfmt == 0 || d->wfmt == 2)),
               "split precision: out.lo missing, or a chain / dx-in-N conv");
    const bool ncat = d->wfmt == 2;        // split precision with N-concatenated weights [tap][Cin/8][2 Cout][8]
    POCO_CHECK(!ncat || (split && out.C <= 64), "weight format 2 needs split precision and Cout <= 64");
    POCO_CHECK(split || out.lo == nullptr, "out.lo given but in.lo is null");
    POCO_CHECK(!split || !d->residual || d->residual_lo, "split precision: residual_lo missing");
    ConvTcParams p{};
    p.n_segs = n_segs;
    p.flags = flags;
    p.in_plane = in.plane_stride;
    p.Hin = in.H; p.Win = in.W;
    p.out_plane = out.plane_stride;
    p.Hout = out.H; p.Wout = out.W;
    p.res_plane = 0;
    bool any_res = false;
    for (int i = 0; i < n_segs; ++i) {
        const poco_conv& c = segs[i];
        p.seg[i].in = static_cast<const __half*>(c.in.data);
        p.seg[i].out = static_cast<__half*>(c.out.data);
        p.seg[i].res = static_cast<const __half*>(c.residual);
        p.seg[i].w = static_cast<const __half*>(c.weight);
        p.seg[i].bias = c.bias;
        p.seg[i].relu = c.relu;
        p.seg[i].in_lo = static_cast<const __half*>(c.in.lo);
        p.seg[i].out_lo = static_cast<__half*>(c.out.lo);
        p.seg[i].res_lo = static_cast<const __half*>(c.residual_lo);
        if (c.residual) {
            POCO_CHECK(!any_res || p.res_plane == c.res_plane_stride, "chain: residual plane strides differ");
            p.res_plane = c.res_plane_stride;
            any_res = true;
        }
        if (i == 0) continue;
        POCO_CHECK(flags != nullptr, "chain: null flags");
        POCO_CHECK(c.in.C == in.C && c.out.C == out.C && c.in.N == in.N && c.in.H == in.H && c.in.W == in.W &&
                       c.out.H == out.H && c.out.W == out.W && c.kh == d->kh && c.kw == d->kw && c.stride == d->stride &&
                       c.pad == d->pad && c.in.plane_stride == in.plane_stride && c.out.plane_stride == out.plane_stride,
                   "chain: segments must share one geometry");
        POCO_CHECK(in.C == out.C, "chain: Cin must equal Cout");
        POCO_CHECK(c.in.data == segs[i - 1].out.data, "chain: segment i must read what segment i-1 wrote");
        // a residual is fetched up to one segment ahead of its use: it may come from outside the chain or
        // from a segment at least two before; and nothing may overwrite a buffer a later segment still reads
        POCO_CHECK(c.residual != segs[i - 1].out.data, "chain: residual must not be the previous segment's output");
        POCO_CHECK(c.out.data != c.in.data && c.out.data != c.residual, "chain: in-place segments are not supported");
    }
    p.Cin = in_s2d ? in.C / 4 : in.C; p.Cout = out.C;         // Cin: channels the weights see
    p.in_s2d = in_s2d ? 1 : 0;
    p.s2d_out = static_cast<__half*>(d->out_s2d.data);
    p.s2d_out_lo = static_cast<__half*>(d->out_s2d.lo);
    p.s2d_plane = d->out_s2d.plane_stride;
    p.s2d_Wp = d->out_s2d.W + 2;
    p.s2d_HpWp = (d->out_s2d.H + 2) * (d->out_s2d.W + 2);
    p.s2d_only = d->s2d_only;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad;
    p.w_bufs = 1;
    {
        const char* dbg = getenv("POCO_CONV_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
        const char* pr = getenv("POCO_CONV_PROF");         // device address (decimal) of 16 zeroed uint64 counters
        p.prof = pr ? reinterpret_cast<unsigned long long*>(strtoull(pr, nullptr, 10)) : nullptr;
        if (p.prof == nullptr) p.debug &= ~(32 | 64);
    }
    p.P_out = int64_t(out.N) * (out.H + 2) * (out.W + 2);
    // weight format 1 = dx-in-N: [3 filter rows][Cin/8][3 * Cout (s-major)][8], see conv_tc_kernel
    const bool dxn = d->wfmt == 1;
    POCO_CHECK(d->wfmt >= 0 && d->wfmt <= 2, "unknown weight format");
    p.tile_stride = dxn ? kTileM - 2 : kTileM;
    p.tile_origin = dxn ? -1 : 0;
    p.num_m_tiles = int((p.P_out + p.tile_stride - 1) / p.tile_stride);

    // N blocking: largest multiple-of-16 divisor of Cout that is <= 256 (dx-in-N: the three column groups)
    int n_tile = 0;
    // (split mode: N <= 128, so that a K chunk of [hi | lo] activations plus [W_hi | W_lo] still leaves two stages)
    int n_cap = split ? 128 : 256;
    // Weight-streamed stride-1 convs (the weights of one N block do not fit next to the operand stages): every tile
    // re-reads the whole weight block from L2, and at ~42 B/clk per SM (6300 B/clk chip-wide L2 cap) that stream -- not
    // the MMAs -- bounded 128->128 @14, 256->256 @7 / @56 (r02 cycle accounting: 103 cycles per MMA against a 64-cycle
    // pipe).  They run with N <= 128 so that four accumulators fit in TMEM and G = 2 tiles share every streamed
    // weight stage (half the L2 traffic; the MMA rate per output column is the same at N = 128 and N = 256).
    static const int stream_group = [] { const char* e = getenv("POCO_B200_STREAM_GROUP"); return e ? atoi(e) : 2; }();
    const bool linear_geom = d->stride == 1 && in.H == out.H && in.W == out.W &&
                             ((d->kh == 3 && d->kw == 3 && d->pad == 1) || (d->kh == 1 && d->kw == 1 && d->pad == 0));
    const bool stream_grouped = stream_group >= 2 && linear_geom && !d->in_s2d && d->wfmt == 0 && n_segs == 1 &&
                                int64_t(d->kh) * d->kw * in.C * std::min(n_cap, out.C) * 2 * sp > (split ? 160 : 112) * 1024;
    if (stream_grouped && out.C % 128 == 0) n_cap = 128;
    if (split && d->in_s2d) n_cap = 64;         // four phase blocks of [hi | lo] planes per K chunk: keep the weight stage small
    for (int t = std::min(n_cap, out.C); t >= 16; t -= 16)
        if (out.C % t == 0) { n_tile = t; break; }
    if (dxn) n_tile = 3 * out.C;
    POCO_CHECK(n_tile > 0 && n_tile <= 256, "no valid N tile");
    p.n_tile = n_tile;
    const int n_blocks = dxn ? 1 : out.C / n_tile;
    int cols = 32;                      // accumulator buffer pitch: power of two >= n_tile (two column groups with N concatenation)
    while (cols < n_tile * (ncat ? 2 : 1)) cols <<= 1;
    // Two CTAs per SM ("half" configuration: <= 110 KB of shared memory, <= 256 TMEM columns) for the layers with
    // N <= 64: measured on one box against the one-CTA configuration, 32->32 @56 33.6 -> 30.3 us (no residual) and
    // 42.6 -> 40.0 us (residual), 64->64 @28 29.5 -> 27.2 us -- two independent producer / issuer / epilogue sets
    // hide each other's barrier round trips, and a finishing CTA's slot is refilled while its neighbour still
    // computes.  POCO_B200_HALF=0 disables it.
    // Only outside plan lanes (max_ctas == 0): a lane's half-size CTAs would spread one per SM and leave the
    // other lanes' full-size CTAs no empty SM (A/B inside the HR modules: 19.5 k vs 19.8 k crops/s).
    static const int half_mode = [] { const char* e = getenv("POCO_B200_HALF"); return e ? atoi(e) : 1; }();   // 0 off, 1 outside lanes, 2 always
    const bool half_stride1 = d->stride == 1 && in.H == out.H && in.W == out.W && d->kh == 3 && d->pad == 1;
    bool half = half_mode > 0 && (half_mode == 2 || d->max_ctas == 0) && !dxn && !split && n_segs == 1 && n_tile <= 64 && half_stride1 &&
                d->out_s2d.data == nullptr;
    int smem_budget = kSmemBudget;
    auto set_half = [&](bool h) {
        half = h;
        smem_budget = h ? kSmemBudgetHalf : kSmemBudget;
        p.acc_bufs = std::max(1, std::min(kMaxAccBufs, (h ? 256 : 512) / cols));      // the MMA warp may run this many tiles ahead of the epilogue
        p.tmem_cols = cols * p.acc_bufs;
    };
    set_half(half);

    const int taps = dxn ? 3 : d->kh * d->kw;       // taps the MMA loop walks
    p.taps = taps;
    // (a stride-2 conv on a phase-split input walks one halo run per plane like a 3x3 / stride 1 conv: linear mode)
    const bool linear = in_s2d || (d->stride == 1 && in.H == out.H && in.W == out.W &&
                                   ((d->kh == 3 && d->kw == 3 && d->pad == 1) || (d->kh == 1 && d->kw == 1 && d->pad == 0)));
    const int mode = linear ? MODE_LINEAR : MODE_GATHER;
    const int cin_w = in_s2d ? in.C / 4 : in.C;         // input channels the weights see
    const int phases = in_s2d ? 4 : 1;
    POCO_CHECK(!dxn || (linear && d->kh == 3 && n_segs == 1 && out.C % 32 == 0 && 3 * out.C <= 256),
               "dx-in-N weights need a single 3x3 / stride 1 / pad 1 conv with 32 or 64 output channels");
    for (int i = 1; i < n_segs; ++i) POCO_CHECK(segs[i].wfmt == 0, "chain: dx-in-N weights are not supported");
    POCO_CHECK(n_segs == 1 || (linear && out.W + 3 <= kTileM), "chain: only 3x3/s1/p1 and 1x1/s1 convolutions chain");
    int budget = 0;         // set per attempt below
    const int w_total = taps * cin_w * n_tile * 2 * sp;
    const int w_res_cap = split ? 160 * 1024 : 112 * 1024;      // largest weight block kept resident

    // M grouping: G adjacent tiles of a 3x3 conv are fetched as ONE run per plane, so the (W+3)-pixel halo is
    // paid once per G tiles (a 128-pixel tile of a 56-wide image reads 246 pixels: 1.9x the data it owns;
    // four tiles read 630 for 512: 1.23x) and one barrier round trip feeds G tiles.  Each tile keeps its own
    // accumulator; 2 G accumulators must fit so that the epilogue still overlaps the next unit's MMAs.
    // Measured at batch 256 (tools/conv_bench.py, POCO_B200_MGROUP=1|2|4): 32->32 @56 36.3 / 47.7 us (no residual /
    // residual) with G = 2 against 39.8 / 54.4 with G = 1 and 36.9 / 50.3 with G = 4; 64->64 @28 and 128->128 @14
    // do not gain (their halo is a smaller share and G > 1 leaves them fewer stages), so the default groups
    // only N <= 32.  POCO_B200_MGROUP=<g> forces up to g everywhere.
    static const int max_group = [] { const char* e = getenv("POCO_B200_MGROUP"); return e ? atoi(e) : 0; }();
    int g_first = 1;
    if (mode == MODE_LINEAR && n_segs == 1 && (taps == 9 || dxn || (taps == 1 && stream_grouped))) {
        const int cap = max_group > 0 ? max_group : (n_tile <= 32 ? 2 : ((stream_grouped && w_total > w_res_cap) ? stream_group : 1));
        while (g_first * 2 <= cap && g_first * 4 <= p.acc_bufs) g_first *= 2;
    }
    auto set_group = [&](int G) {
        p.m_group = G;
        p.halo = dxn ? (out.W + 2) : (taps == 9 ? (out.W + 2) + 1 : 0);
        p.halo_hi = in_s2d ? 0 : p.halo;        // the taps of a phase-split input reach back only
        p.a_copy_bytes = ((G - 1) * p.tile_stride + kTileM + p.halo + p.halo_hi) * 16;
        p.a_plane_bytes = round_up(p.a_copy_bytes, 128);
    };
    if (mode == MODE_LINEAR) {
        set_group(g_first);
    } else {
        p.m_group = 1;
        p.halo = 0;
        p.a_copy_bytes = kTileM * 16;
        p.a_plane_bytes = kTileM * 16;
    }
    // choose epilogue item width, K chunk, residency and stage count.  Wide items (32 columns) halve the
    // per-item synchronisation of the epilogue; they are used when N >= 64 and the rings still leave
    // room for >= 4 operand stages; only the 1x1 convs (small K, epilogue bound) take them -- for every
    // other shape the operand stages are the better use of shared memory (measured, profiles/).
    bool found = false;
    p.tap_group = 1;
    for (int attempt = half ? 0 : 1; attempt < 2 && !found; ++attempt) {       // the half configuration first, when eligible
    set_half(attempt == 0);
    for (int G = std::min(p.acc_bufs >= 4 ? g_first : 1, g_first); G >= 1 && !found; G >>= 1) {
    if (mode == MODE_LINEAR) set_group(G);
    for (int want4 = (mode == MODE_LINEAR && (dxn || (n_tile >= 64 && taps == 1) || n_tile == 32) ? 1 : 0); want4 >= (dxn ? 1 : 0) && !found; --want4) {
        p.item_planes = want4 ? 4 : 2;
        const int items_here = ((dxn ? out.C : n_tile) + p.item_planes * 8 - 1) / (p.item_planes * 8);
        budget = smem_budget - kHeaderBytes - ring_bytes_for(p.item_planes, half || split, sp, any_res);
        (void)items_here;
        if (mode == MODE_LINEAR) {
            const int kcs[4] = {64, 48, 32, 16};
            // resident weights: a chain double-buffers them when that still leaves >= 2 operand stages
            static const int max_wbufs = [] { const char* e = getenv("POCO_CHAIN_WBUFS"); return e ? atoi(e) : 2; }();
            for (int wb = (n_segs > 1 ? std::min(2, max_wbufs) : 1); wb >= 0 && !found; --wb) {
                const int resident = wb > 0;
                if (resident && w_total > w_res_cap) continue;
                for (int ki = 0; ki < 4 && !found; ++ki) {
                    const int kc = kcs[ki];
                    if (cin_w % kc != 0) continue;
                    const int a_stage = (kc / 8) * phases * p.a_plane_bytes * sp;
                    const int w_stage = resident ? 0 : taps * kc * n_tile * 2 * sp;
                    const int avail = budget - wb * w_total;
                    const int stages = std::min(max_stages(), avail / (a_stage + w_stage));
                    if (stages < (want4 ? 4 : 2)) continue;
                    p.kc = kc; p.n_chunks = cin_w / kc;
                    p.w_resident = resident;
                    p.w_bufs = std::max(1, wb);
                    p.w_res_bytes = resident ? w_total : 0;
                    p.a_stage_bytes = a_stage; p.w_stage_bytes = w_stage;
                    static const int max_rings = [] { const char* e = getenv("POCO_B200_RINGS"); return e ? atoi(e) : 2; }();
                    p.rings = (stages >= 4 && max_rings >= 2) ? 2 : 1;
                    p.stages = stages / p.rings;
                    found = true;
                }
            }
        } else {
            // gather: K chunk = 16 channels (2 planes); a stage holds `tap_group` taps (all / one row / one)
            const int groups[3] = {taps <= 9 ? taps : d->kw, d->kw, 1};
            for (int resident = 1; resident >= 0 && !found; --resident) {
                if (resident && w_total > w_res_cap) continue;
                for (int gi = 0; gi < 3 && !found; ++gi) {
                    const int tg = groups[gi];
                    if (taps % tg != 0) continue;
                    const int a_stage = tg * 4096 * sp;
                    const int w_stage = resident ? 0 : tg * 16 * n_tile * 2 * sp;
                    const int avail = budget - (resident ? w_total : 0);
                    const int stages = std::min(max_stages(), avail / (a_stage + w_stage));
                    if (stages < (want4 ? 4 : kGatherLag + 1)) continue;
                    p.kc = 16; p.n_chunks = in.C / 16;
                    p.tap_group = tg;
                    p.w_resident = resident;
                    p.w_res_bytes = resident ? w_total : 0;
                    p.a_stage_bytes = a_stage; p.w_stage_bytes = w_stage;
                    p.rings = 1;
                    p.stages = stages;
                    found = true;
                }
            }
        }
    }
    }
    }
    POCO_CHECK(found, "no shared-memory configuration fits this convolution");
    // every epilogue warp that owns a share of a tile signals it once (per N block)
    p.n_out = dxn ? out.C : n_tile;
    p.epi_items = (p.n_out + p.item_planes * 8 - 1) / (p.item_planes * 8);
    p.flag_expect = (p.epi_items >= 2 ? 8 : 4) * n_blocks;
    // slabs are n_tile*16 bytes (a multiple of 256): the resident region needs no padding and
    // w_res_bytes is both the region size and the mbarrier transaction count
    size_t smem = size_t(kHeaderBytes) + ring_bytes_for(p.item_planes, half || split, sp, any_res) + size_t(p.w_res_bytes) * p.w_bufs + size_t(p.stages * p.rings) * (p.a_stage_bytes + p.w_stage_bytes);
    p.res_ring = (any_res || ring_always()) ? res_ring_base(half || split) : 0;
    if (any_res && p.res_ring > 0) {       // spend spare shared memory on a deeper residual prefetch ring
        const int item_bytes = p.item_planes * kPlaneTile * sp;
        const int extra = int((size_t(smem_budget) - smem) / (2 * item_bytes));
        const int base = p.res_ring;
        p.res_ring = std::min(kMaxResRing, base + std::max(0, extra));
        smem += size_t(p.res_ring - base) * 2 * item_bytes;
    }
    const int sm_budget = (half ? 2 : 1) * (d->max_ctas > 0 ? std::min(d->max_ctas, num_sms()) : num_sms());
    const int num_units = (p.num_m_tiles + p.m_group - 1) / p.m_group;
    dim3 grid(std::max(1, std::min(num_units, sm_budget / n_blocks)), n_blocks);
    const ConvTcParams& pk = p;
    // Programmatic dependent launch is opt-in (POCO_B200_PDL=1): one-CTA-per-SM kernels leave the dependent grid no
    // room to start early, and its CTAs parked in griddepcontrol.wait cost 1.7 % end to end (A/B: 20.0 k with,
    // 20.4 k crops/s without; round 2: 12.63 vs 12.32 ms).  Round 2 also tried the co-resident variant -- every
    // stride-1 conv as ONE half-size CTA per SM (<= 110 KB, <= 256 TMEM columns) with PDL, so that the successor's
    // CTAs sit set up next to the running ones: the overlap works (the successor's first data lands 2 us after the
    // predecessor's exit instead of 7, profiles/r02b_conv_launch_timeline.csv) but half the shared memory per conv
    // costs far more than the overlap returns (20.6 vs 12.3 ms per step, profiles/r02b_coop_pdl_experiment.csv).
    static const bool use_pdl = [] { const char* e = getenv("POCO_B200_PDL"); return e && e[0] == '1'; }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    auto launch = [&](auto kernel, std::once_flag& flag, int threads) -> cudaError_t {
        std::call_once(flag, [&] { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget); });
        cfg.blockDim = dim3(threads);
        return cudaLaunchKernelEx(&cfg, kernel, pk);
    };
    static std::once_flag once4[13], once_s2d[6];
    POCO_CHECK(!ncat || mode == MODE_LINEAR, "weight format 2 is for stride-1 3x3 / 1x1 convs");
    const bool s2dout = p.s2d_out != nullptr;
    POCO_CHECK(!s2dout || (mode == MODE_LINEAR && !dxn), "out_s2d needs a conv on the halo-run path");
    if (s2dout) {       // (never the two-CTA "half" flavour: set below)
        const int sv = ncat ? 2 : (split ? 1 : 0);
        if (sv == 0 && p.item_planes == 2) POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false, 1, 0, true>, once_s2d[0], Roles<MODE_LINEAR>::kThreads));
        else if (sv == 0) POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false, 1, 0, true>, once_s2d[1], Roles<MODE_LINEAR>::kThreads));
        else if (sv == 1 && p.item_planes == 2) POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false, 1, 1, true>, once_s2d[2], Roles<MODE_LINEAR>::kThreads));
        else if (sv == 1) POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false, 1, 1, true>, once_s2d[3], Roles<MODE_LINEAR>::kThreads));
        else if (p.item_planes == 2) POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false, 1, 2, true>, once_s2d[4], Roles<MODE_LINEAR>::kThreads));
        else POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false, 1, 2, true>, once_s2d[5], Roles<MODE_LINEAR>::kThreads));
    } else if (ncat && p.item_planes == 2)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false, 1, 2>, once4[11], Roles<MODE_LINEAR>::kThreads));
    else if (ncat)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false, 1, 2>, once4[12], Roles<MODE_LINEAR>::kThreads));
    else if (split && mode == MODE_LINEAR && p.item_planes == 2)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false, 1, 1>, once4[7], Roles<MODE_LINEAR>::kThreads));
    else if (split && mode == MODE_LINEAR)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false, 1, 1>, once4[8], Roles<MODE_LINEAR>::kThreads));
    else if (split && p.item_planes == 2)
        POCO_CUDA(launch(conv_tc_kernel<MODE_GATHER, 2, false, 1, 1>, once4[9], Roles<MODE_GATHER>::kThreads));
    else if (split)
        POCO_CUDA(launch(conv_tc_kernel<MODE_GATHER, 4, false, 1, 1>, once4[10], Roles<MODE_GATHER>::kThreads));
    else if (half && p.item_planes == 2)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false, 2>, once4[5], Roles<MODE_LINEAR>::kThreads));
    else if (half)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false, 2>, once4[6], Roles<MODE_LINEAR>::kThreads));
    else if (dxn)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, true>, once4[4], Roles<MODE_LINEAR>::kThreads));
    else if (mode == MODE_LINEAR && p.item_planes == 2)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 2, false>, once4[0], Roles<MODE_LINEAR>::kThreads));
    else if (mode == MODE_LINEAR)
        POCO_CUDA(launch(conv_tc_kernel<MODE_LINEAR, 4, false>, once4[1], Roles<MODE_LINEAR>::kThreads));
    else if (p.item_planes == 2)
        POCO_CUDA(launch(conv_tc_kernel<MODE_GATHER, 2, false>, once4[2], Roles<MODE_GATHER>::kThreads));
    else
        POCO_CUDA(launch(conv_tc_kernel<MODE_GATHER, 4, false>, once4[3], Roles<MODE_GATHER>::kThreads));
    POCO_LAUNCHED();
    return 0;
}

}  // namespace poco
