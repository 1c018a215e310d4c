// poco_b200 -- head kernels (fp32 math): small-M linear layers, rot6d -> rotmat, the fused PARE
// part-attention head and the conditional RealNVP passes.  These are <0.1 % of the FLOPs of a
// forward pass; they are written for correctness-first fp32 parity and few launches.
#include <algorithm>

#include <mutex>

#include "common.cuh"
#include "internal.h"

namespace poco {

namespace {

// ------------------------------------------------------------------------------------------------
// y = act(x W^T + b) + res      x [M,I] (ldx), W [O,I] row-major, y [M,O] (ldy)
// ------------------------------------------------------------------------------------------------
constexpr int LBM = 32, LBN = 32, LBK = 32;

// 32x32 output tile, K step 32, 256 threads (2x2 outputs each); the next K tile is prefetched into
// registers while the current one is multiplied, so a CTA has two tiles of loads in flight.
// blockIdx.z = K split: with `splits` > 1 every CTA multiplies one K range and writes its partial sums to
// d.scratch [split][M][O]; linear_reduce_kernel then adds them up and applies bias / activation / residual.
__global__ void __launch_bounds__(256) linear_kernel(poco_linear d, int splits, int k_per_split) {
    __shared__ float xs[LBK][LBM + 1];
    __shared__ float ws[LBK][LBN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx -> O, ty -> M
    const int m0 = blockIdx.y * LBM, o0 = blockIdx.x * LBN;
    const int k_begin = blockIdx.z * k_per_split, k_end = min(d.I, k_begin + k_per_split);
    // each thread stages 4 x elements and 4 w elements per K tile: element e = threadIdx.x + 256*i
    const int lk = threadIdx.x & 31, lr = threadIdx.x >> 5;     // k within tile, row group (8 groups x 4 rows)
    float px[4], pw[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lr + 8 * i, k = k0 + lk;
            px[i] = (m0 + r < d.M && k < k_end) ? d.x[(long long)(m0 + r) * d.ldx + k] : 0.f;
            pw[i] = (o0 + r < d.O && k < k_end) ? d.w[(long long)(o0 + r) * d.I + k] : 0.f;
        }
    };
    float acc[2][2] = {};
    fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += LBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xs[lk][lr + 8 * i] = px[i];
            ws[lk][lr + 8 * i] = pw[i];
        }
        __syncthreads();
        if (k0 + LBK < k_end) fetch(k0 + LBK);
#pragma unroll
        for (int k = 0; k < LBK; ++k) {
            const float a0 = xs[k][ty], a1 = xs[k][ty + 16];
            const float b0 = ws[k][tx], b1 = ws[k][tx + 16];
            acc[0][0] = fmaf(a0, b0, acc[0][0]);
            acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]);
            acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + ty + 16 * i, o = o0 + tx + 16 * j;
            if (m < d.M && o < d.O) {
                if (splits > 1) {
                    d.scratch[((long long)blockIdx.z * d.M + m) * d.O + o] = acc[i][j];
                    continue;
                }
                float v = acc[i][j] + (d.b ? d.b[o] : 0.f);
                if (d.act == 1) v = 1.f / (1.f + expf(-v));
                else if (d.act == 2) v = v > 20.f ? v : log1pf(expf(v));      // nn.Softplus (beta=1, threshold=20)
                if (d.res) v += d.res[(long long)m * d.ldres + o];
                d.y[(long long)m * d.ldy + o] = v;
            }
        }
}

__global__ void __launch_bounds__(256) linear_reduce_kernel(poco_linear d, int splits) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d.M * d.O) return;
    const int m = idx / d.O, o = idx - m * d.O;
    float v = d.b ? d.b[o] : 0.f;
    for (int s = 0; s < splits; ++s) v += d.scratch[((long long)s * d.M + m) * d.O + o];
    if (d.act == 1) v = 1.f / (1.f + expf(-v));
    else if (d.act == 2) v = v > 20.f ? v : log1pf(expf(v));
    if (d.res) v += d.res[(long long)m * d.ldres + o];
    d.y[(long long)m * d.ldy + o] = v;
}

__global__ void copy2d_kernel(poco_copy2d d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)d.rows * d.cols) return;
    const int r = int(idx / d.cols), c = int(idx % d.cols);
    d.dst[(long long)r * d.ldd + c] = d.src[(long long)(d.bcast ? 0 : r) * d.lds + c];
}

// geometry.py:247-261 -- the 6 numbers are a 3x2 row-major matrix: a1 = (x0,x2,x4), a2 = (x1,x3,x5)
__device__ __forceinline__ void rot6d_one(const float* x, float* R) {
    const float a1[3] = {x[0], x[2], x[4]}, a2[3] = {x[1], x[3], x[5]};
    const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    const float dp = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const float u[3] = {a2[0] - dp * b1[0], a2[1] - dp * b1[1], a2[2] - dp * b1[2]};
    const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        R[3 * i + 0] = b1[i];
        R[3 * i + 1] = b2[i];
        R[3 * i + 2] = b3[i];
    }
}

__global__ void rot6d_kernel(poco_rot6d d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const float* x = d.x + (long long)(i / d.per_row) * d.ldx + (i % d.per_row) * 6;
    float v[6], R[9];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = x[k];
    rot6d_one(v, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) d.out[(long long)i * 9 + k] = R[k];
}

// ------------------------------------------------------------------------------------------------
// PARE head
// ------------------------------------------------------------------------------------------------
constexpr int PJ = 24, PC = 128, PS = 64, PCHUNK = 128;

__device__ __forceinline__ void unpack8h(uint4 v, float* f) {
    float2 t;
    t = unpack_half2(v.x); f[0] = t.x; f[1] = t.y;
    t = unpack_half2(v.y); f[2] = t.x; f[3] = t.y;
    t = unpack_half2(v.z); f[4] = t.x; f[5] = t.y;
    t = unpack_half2(v.w); f[6] = t.x; f[7] = t.y;
}

// K1: segm[n, j, y, x] = b_kp[j] + sum_c w_kp[j, c] part_feats[n, c, y, x]   (fp32, from fp16 feats)
__global__ void __launch_bounds__(128) pare_logits_kernel(poco_pare_head d) {
    __shared__ float w[25 * PC];
    __shared__ float b[25];
    for (int i = threadIdx.x; i < 25 * PC; i += blockDim.x) w[i] = d.w_kp[i];
    if (threadIdx.x < 25) b[threadIdx.x] = d.b_kp[threadIdx.x];
    __syncthreads();
    const int H = d.part_feats.H, W = d.part_feats.W;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)d.part_feats.N * H * W) return;
    const int x = int(idx % W), y = int((idx / W) % H), n = int(idx / ((long long)W * H));
    const long long pix = (long long)n * (H + 2) * (W + 2) + (long long)(y + 1) * (W + 2) + (x + 1);
    const __half* base = static_cast<const __half*>(d.part_feats.data);
    const __half* base_lo = static_cast<const __half*>(d.part_feats.lo);       // split-precision mode: value = hi + lo
    float f[PC];
#pragma unroll
    for (int pl = 0; pl < PC / 8; ++pl)
        unpack8h(*reinterpret_cast<const uint4*>(base + ((long long)pl * d.part_feats.plane_stride + pix) * 8), f + pl * 8);
    if (base_lo != nullptr) {
#pragma unroll
        for (int pl = 0; pl < PC / 8; ++pl) {
            float l[8];
            unpack8h(*reinterpret_cast<const uint4*>(base_lo + ((long long)pl * d.part_feats.plane_stride + pix) * 8), l);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[pl * 8 + i] += l[i];
        }
    }
    const long long hw = (long long)H * W;
    for (int j = 0; j < 25; ++j) {
        float acc = b[j];
        const float4* wj = reinterpret_cast<const float4*>(w + j * PC);
#pragma unroll
        for (int c = 0; c < PC / 4; ++c) {
            const float4 ww = wj[c];
            acc = fmaf(f[4 * c], ww.x, acc);
            acc = fmaf(f[4 * c + 1], ww.y, acc);
            acc = fmaf(f[4 * c + 2], ww.z, acc);
            acc = fmaf(f[4 * c + 3], ww.w, acc);
        }
        d.segm[((long long)n * 25 + j) * hw + (long long)y * W + x] = acc;
    }
}

// scratch layout per (crop, chunk): [PJ] max, [PJ] sum, [PJ][PC] pooled
__host__ __device__ inline int pare_chunk_rows(int H, int W) { return max(1, min(H, PCHUNK / W)); }
__host__ __device__ inline int pare_num_chunks(int H, int W) { const int r = pare_chunk_rows(H, W); return (H + r - 1) / r; }
constexpr int kPartial = PJ + PJ + PJ * PC;

// K2: flash-style partial softmax pooling over one chunk of rows of one crop
// SPLIT (split-precision mode): the features are hi + lo fp16 pairs and are staged as fp32 in dynamic shared memory
template <bool SPLIT>
__global__ void __launch_bounds__(256) pare_pool_kernel(poco_pare_head d) {
    __shared__ float a[PCHUNK][PJ];
    __shared__ __align__(16) __half fs[SPLIT ? 1 : PCHUNK][PC + 8];
    extern __shared__ __align__(16) float fsf_raw[];
    float(*fsf)[PC + 4] = reinterpret_cast<float(*)[PC + 4]>(fsf_raw);      // [PCHUNK][PC + 4], SPLIT only
    __shared__ float mloc[PJ];
    const int H = d.smpl_feats.H, W = d.smpl_feats.W;
    const int rows = pare_chunk_rows(H, W), nch = pare_num_chunks(H, W);
    const int n = blockIdx.x / nch, ch = blockIdx.x % nch;
    const int y0 = ch * rows, y1 = min(H, y0 + rows);
    const int npx = (y1 - y0) * W;
    const long long hw = (long long)H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // attention logits (joints 1..24 of segm; channel 0 is background, pare_head.py:796)
    for (int e = threadIdx.x; e < PJ * npx; e += blockDim.x) {
        const int j = e / npx, px = e % npx;
        a[px][j] = d.segm[((long long)n * 25 + 1 + j) * hw + (long long)y0 * W + px];
    }
    // smpl features of the chunk -> shared (coalesced 16-byte loads along a plane)
    const __half* base = static_cast<const __half*>(d.smpl_feats.data);
    for (int e = threadIdx.x; e < (PC / 8) * npx; e += blockDim.x) {
        const int pl = e / npx, px = e % npx;
        const int y = y0 + px / W, x = px % W;
        const long long pix = (long long)n * (H + 2) * (W + 2) + (long long)(y + 1) * (W + 2) + (x + 1);
        const uint4 hv = *reinterpret_cast<const uint4*>(base + ((long long)pl * d.smpl_feats.plane_stride + pix) * 8);
        if (SPLIT) {
            const uint4 lv = *reinterpret_cast<const uint4*>(static_cast<const __half*>(d.smpl_feats.lo) +
                                                             ((long long)pl * d.smpl_feats.plane_stride + pix) * 8);
            float h[8], l[8];
            unpack8h(hv, h);
            unpack8h(lv, l);
            *reinterpret_cast<float4*>(&fsf[px][pl * 8]) = make_float4(h[0] + l[0], h[1] + l[1], h[2] + l[2], h[3] + l[3]);
            *reinterpret_cast<float4*>(&fsf[px][pl * 8 + 4]) = make_float4(h[4] + l[4], h[5] + l[5], h[6] + l[6], h[7] + l[7]);
        } else {
            *reinterpret_cast<uint4*>(&fs[px][pl * 8]) = hv;
        }
    }
    __syncthreads();
    float* part = d.scratch + (long long)blockIdx.x * kPartial;
    for (int j = warp; j < PJ; j += 8) {
        float m = -INFINITY;
        for (int px = lane; px < npx; px += 32) m = fmaxf(m, a[px][j]);
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) { mloc[j] = m; part[j] = m; }
    }
    __syncthreads();
    for (int j = warp; j < PJ; j += 8) {
        float s = 0.f;
        for (int px = lane; px < npx; px += 32) {
            const float e = expf(a[px][j] - mloc[j]);
            a[px][j] = e;
            s += e;
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) part[PJ + j] = s;
    }
    __syncthreads();
    // thread -> 4 channels x 3 joints
    const int cg = threadIdx.x & 31, jg = threadIdx.x >> 5;
    float acc[3][4] = {};
    for (int px = 0; px < npx; ++px) {
        float2 f01, f23;
        if (SPLIT) {
            const float4 v4 = *reinterpret_cast<const float4*>(&fsf[px][cg * 4]);
            f01 = make_float2(v4.x, v4.y);
            f23 = make_float2(v4.z, v4.w);
        } else {
            const uint2 raw = *reinterpret_cast<const uint2*>(&fs[px][cg * 4]);
            f01 = unpack_half2(raw.x);
            f23 = unpack_half2(raw.y);
        }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const float w = a[px][jg * 3 + jj];
            acc[jj][0] = fmaf(w, f01.x, acc[jj][0]);
            acc[jj][1] = fmaf(w, f01.y, acc[jj][1]);
            acc[jj][2] = fmaf(w, f23.x, acc[jj][2]);
            acc[jj][3] = fmaf(w, f23.y, acc[jj][3]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < 3; ++jj)
#pragma unroll
        for (int k = 0; k < 4; ++k) part[2 * PJ + (jg * 3 + jj) * PC + cg * 4 + k] = acc[jj][k];
}

// K3: combine partials, per-joint MLPs, shape / cam linears, rot6d.  One CTA per crop.
__global__ void __launch_bounds__(256) pare_final_kernel(poco_pare_head d) {
    __shared__ float pl[PC][PJ + 1];      // point_local[c][j]
    __shared__ float cs[PS * PJ];         // cam_shape flattened c*24 + j
    __shared__ float scale[32][PJ];       // per-chunk rescale (nch <= 32)
    __shared__ float denom[PJ];
    __shared__ float p6[PJ][6];
    const int H = d.smpl_feats.H, W = d.smpl_feats.W;
    const int nch = pare_num_chunks(H, W);
    const int n = blockIdx.x;
    const float* part = d.scratch + (long long)n * nch * kPartial;
    if (threadIdx.x < PJ) {
        const int j = threadIdx.x;
        float M = -INFINITY;
        for (int k = 0; k < nch; ++k) M = fmaxf(M, part[(long long)k * kPartial + j]);
        float s = 0.f;
        for (int k = 0; k < nch; ++k) {
            const float sc = expf(part[(long long)k * kPartial + j] - M);
            scale[k][j] = sc;
            s += sc * part[(long long)k * kPartial + PJ + j];
        }
        denom[j] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < PJ * PC; e += blockDim.x) {
        const int j = e / PC, c = e % PC;
        float v = 0.f;
        for (int k = 0; k < nch; ++k) v = fmaf(part[(long long)k * kPartial + 2 * PJ + e], scale[k][j], v);
        v /= denom[j];
        pl[c][j] = v;
        d.uncert_feat[(long long)n * (PC * PJ) + c * PJ + j] = v;       // reshape of [B,128,24]
    }
    __syncthreads();
    // cam_shape[o][j] = b_sf[o] + sum_c w_sf[o][c] pl[c][j]  (1x1 smpl_final_layer after the pooling)
    for (int e = threadIdx.x; e < PS * PJ; e += blockDim.x) {
        const int o = e / PJ, j = e % PJ;
        float v = d.b_sf[o];
        for (int c = 0; c < PC; ++c) v = fmaf(d.w_sf[o * PC + c], pl[c][j], v);
        cs[e] = v;
    }
    // locally connected pose MLP: pose6[o][j] = sum_c pl[c][j] w_pose[o][c][j]   (locallyconnected2d.py:27-37)
    for (int e = threadIdx.x; e < 6 * PJ; e += blockDim.x) {
        const int o = e / PJ, j = e % PJ;
        float v = 0.f;
        for (int c = 0; c < PC; ++c) v = fmaf(pl[c][j], d.w_pose[(o * PC + c) * PJ + j], v);
        p6[j][o] = v;
        d.pose6d[((long long)n * PJ + j) * 6 + o] = v;
    }
    __syncthreads();
    // shape (10) and cam (3): one warp per output, 1536-long dot products
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < 13; o += 8) {
        const float* wrow = o < 10 ? d.w_shape + o * (PS * PJ) : d.w_cam + (o - 10) * (PS * PJ);
        float v = 0.f;
        for (int i = lane; i < PS * PJ; i += 32) v = fmaf(wrow[i], cs[i], v);
        for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
        if (lane == 0) {
            if (o < 10) d.shape[(long long)n * 10 + o] = v + d.b_shape[o];
            else d.cam[(long long)n * 3 + (o - 10)] = v + d.b_cam[o - 10];
        }
    }
    if (threadIdx.x < PJ) {
        float R[9];
        rot6d_one(p6[threadIdx.x], R);
        for (int k = 0; k < 9; ++k) d.rotmat[((long long)n * PJ + threadIdx.x) * 9 + k] = R[k];
    }
}

// ------------------------------------------------------------------------------------------------
// conditional RealNVP  (real_nvp.py:25-65).  One CTA = RB rows, 128 threads (4 warps).
// ------------------------------------------------------------------------------------------------
constexpr int RB = 8, RMAXD = 16, RMAXH = 64, RMAXIN = 16 + 1024;

__global__ void __launch_bounds__(128) realnvp_kernel(poco_realnvp d) {
    __shared__ float inp[RB][RMAXIN];        // [z * mask, ctx]
    __shared__ float z[RB][RMAXD];
    __shared__ float h0[RB][2 * RMAXH];      // s-net units then t-net units
    __shared__ float h1[RB][2 * RMAXH];
    __shared__ float st[RB][2 * RMAXD];      // s then t
    __shared__ float logdet[RB];
    const int D = d.D, CTX = d.CTX, HID = d.HID, IN = D + CTX;
    const int r0 = blockIdx.x * RB;
    const int nr = min(RB, d.R - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long net_sz = (long long)HID * IN + HID + (long long)HID * HID + HID + (long long)D * HID + D;
    const long long layer_sz = D + 2 * net_sz;
    for (int e = threadIdx.x; e < RB * D; e += blockDim.x) {
        const int r = e / D, k = e % D;
        z[r][k] = r < nr ? d.x[(long long)(r0 + r) * D + k] : 0.f;
    }
    const bool hoisted = d.ctx_part != nullptr;     // the context part of every first layer comes precomputed (one GEMM)
    if (!hoisted)
        for (int e = threadIdx.x; e < RB * CTX; e += blockDim.x) {
            const int r = e / CTX, k = e % CTX;
            inp[r][D + k] = r < nr ? d.ctx[(long long)(r0 + r) * CTX + k] : 0.f;
        }
    if (threadIdx.x < RB) logdet[threadIdx.x] = 0.f;
    __syncthreads();
    for (int step = 0; step < d.L; ++step) {
        const int li = d.direction == 0 ? d.L - 1 - step : step;      // log_prob walks the layers backwards
        const float* P = d.params + li * layer_sz;
        const float* mask = P;
        for (int e = threadIdx.x; e < RB * D; e += blockDim.x) inp[e / D][e % D] = z[e / D][e % D] * mask[e % D];
        __syncthreads();
        // layer 0: 2*HID units, warp-cooperative dot products over IN (coalesced weight reads)
        if (hoisted) {      // D-column part only (D <= 16: one thread per (row, unit)) + the precomputed context part
            for (int e = threadIdx.x; e < RB * 2 * HID; e += blockDim.x) {
                const int r = e / (2 * HID), u = e % (2 * HID);
                const float* net = P + D + (u / HID) * net_sz;
                const float* wrow = net + (long long)(u % HID) * IN;
                const int g = min(r0 + r, d.R - 1) / d.ctx_group;
                float v = d.ctx_part[(long long)g * (d.L * 2 * HID) + (long long)li * 2 * HID + u];      // includes b0
                for (int i = 0; i < D; ++i) v = fmaf(wrow[i], inp[r][i], v);
                h0[r][u] = v > 0.f ? v : 0.01f * v;
            }
        } else
        for (int u = warp; u < 2 * HID; u += 4) {
            const float* net = P + D + (u / HID) * net_sz;
            const float* wrow = net + (long long)(u % HID) * IN;
            float acc[RB] = {};
            for (int i = lane; i < IN; i += 32) {
                const float w = wrow[i];
#pragma unroll
                for (int r = 0; r < RB; ++r) acc[r] = fmaf(w, inp[r][i], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                float v = acc[r];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) {
                    v += net[(long long)HID * IN + (u % HID)];
                    h0[r][u] = v > 0.f ? v : 0.01f * v;
                }
            }
        }
        __syncthreads();
        // layer 1: HID x HID per net
        for (int e = threadIdx.x; e < RB * 2 * HID; e += blockDim.x) {
            const int r = e / (2 * HID), u = e % (2 * HID);
            const float* net = P + D + (u / HID) * net_sz;
            const float* w1 = net + (long long)HID * IN + HID + (long long)(u % HID) * HID;
            float v = net[(long long)HID * IN + HID + (long long)HID * HID + (u % HID)];
            const float* hin = &h0[r][(u / HID) * HID];
            for (int i = 0; i < HID; ++i) v = fmaf(w1[i], hin[i], v);
            h1[r][u] = v > 0.f ? v : 0.01f * v;
        }
        __syncthreads();
        // layer 2: D outputs per net; tanh on s
        for (int e = threadIdx.x; e < RB * 2 * D; e += blockDim.x) {
            const int r = e / (2 * D), k = e % (2 * D);
            const int which = k / D, o = k % D;
            const float* net = P + D + which * net_sz;
            const float* w2 = net + (long long)HID * IN + HID + (long long)HID * HID + HID + (long long)o * HID;
            float v = net[(long long)HID * IN + HID + (long long)HID * HID + HID + (long long)D * HID + o];
            const float* hin = &h1[r][which * HID];
            for (int i = 0; i < HID; ++i) v = fmaf(w2[i], hin[i], v);
            st[r][which * RMAXD + o] = which == 0 ? tanhf(v) : v;
        }
        __syncthreads();
        if (threadIdx.x < RB) {
            const int r = threadIdx.x;
            float ld = 0.f;
            for (int k = 0; k < D; ++k) {
                const float m = mask[k];
                const float s = st[r][k] * (1.f - m), t = st[r][RMAXD + k] * (1.f - m);
                if (d.direction == 0) {
                    z[r][k] = (1.f - m) * (z[r][k] - t) * expf(-s) + z[r][k] * m;
                    ld -= s;
                } else {
                    z[r][k] = z[r][k] * m + (1.f - m) * (z[r][k] * expf(s) + t);
                }
            }
            logdet[r] += ld;
        }
        __syncthreads();
    }
    if (threadIdx.x < nr) {
        const int r = threadIdx.x;
        if (d.direction == 0) {
            float ss = 0.f;
            for (int k = 0; k < D; ++k) ss += z[r][k] * z[r][k];
            d.out[r0 + r] = -0.5f * ss - 0.5f * D * 1.8378770664093453f + logdet[r];   // log(2*pi)
            if (d.logdet_out) d.logdet_out[r0 + r] = logdet[r];
            if (d.z_out)
                for (int k = 0; k < D; ++k) d.z_out[(long long)(r0 + r) * D + k] = z[r][k];
        } else {
            for (int k = 0; k < D; ++k) d.out[(long long)(r0 + r) * D + k] = z[r][k];
        }
    }
}

inline unsigned blocks_for(long long n, int bs) { return unsigned((n + bs - 1) / bs); }

}  // namespace
}  // namespace poco

using namespace poco;

extern "C" int poco_linear_run(const poco_linear* d, void* stream) {
    POCO_CHECK(d->x && d->w && d->y, "null pointer");
    POCO_CHECK(d->M > 0 && d->I > 0 && d->O > 0, "empty problem");
    POCO_CHECK(d->ldx >= d->I && d->ldy >= d->O && (!d->res || d->ldres >= d->O), "leading dimension too small");
    dim3 grid((d->O + LBN - 1) / LBN, (d->M + LBM - 1) / LBM);
    // split K when the output tiles alone cannot fill the GPU: aim at ~4 CTAs per SM, >= 4 K tiles per split
    int splits = 1;
    if (d->scratch != nullptr) {
        const int ctas = int(grid.x * grid.y), ktiles = (d->I + LBK - 1) / LBK;
        splits = std::max(1, std::min(std::min(592 / std::max(1, ctas), ktiles / 4), 8));
        splits = int(std::min<int64_t>(splits, d->scratch_floats / (int64_t(d->M) * d->O)));
        if (splits < 2) splits = 1;
    }
    const int k_per_split = ((d->I + splits - 1) / splits + LBK - 1) / LBK * LBK;
    grid.z = splits;
    linear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, splits, k_per_split);
    POCO_LAUNCHED();
    if (splits > 1) {
        linear_reduce_kernel<<<(d->M * d->O + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, splits);
        POCO_LAUNCHED();
    }
    return 0;
}

extern "C" int poco_copy2d_run(const poco_copy2d* d, void* stream) {
    POCO_CHECK(d->src && d->dst && d->rows > 0 && d->cols > 0, "bad arguments");
    copy2d_kernel<<<blocks_for((long long)d->rows * d->cols, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(*d);
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_rot6d_run(const poco_rot6d* d, void* stream) {
    POCO_CHECK(d->x && d->out && d->n > 0 && d->per_row > 0, "bad arguments");
    rot6d_kernel<<<blocks_for(d->n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(*d);
    POCO_LAUNCHED();
    return 0;
}

extern "C" int64_t poco_pare_scratch_floats(int32_t N, int32_t H, int32_t W) {
    return int64_t(N) * pare_num_chunks(H, W) * kPartial;
}

extern "C" int poco_pare_head_run(const poco_pare_head* d, void* stream) {
    if (check_act(d->part_feats, "part_feats") || check_act(d->smpl_feats, "smpl_feats")) return 1;
    POCO_CHECK(d->part_feats.C == PC && d->smpl_feats.C == PC, "PARE branches must have 128 channels");
    POCO_CHECK(d->part_feats.H == d->smpl_feats.H && d->part_feats.W == d->smpl_feats.W &&
                   d->part_feats.N == d->smpl_feats.N, "branch geometry mismatch");
    const int H = d->part_feats.H, W = d->part_feats.W, N = d->part_feats.N;
    POCO_CHECK(W <= PCHUNK && pare_num_chunks(H, W) <= 32, "feature map too large for the pooling kernel");
    POCO_CHECK(d->segm && d->uncert_feat && d->pose6d && d->rotmat && d->shape && d->cam && d->scratch, "null output");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    pare_logits_kernel<<<blocks_for((long long)N * H * W, 128), 128, 0, s>>>(*d);
    POCO_LAUNCHED();
    POCO_CHECK((d->part_feats.lo != nullptr) == (d->smpl_feats.lo != nullptr), "pare_head: both feature tensors must share one precision mode");
    if (d->smpl_feats.lo != nullptr) {
        constexpr int kDyn = PCHUNK * (PC + 4) * int(sizeof(float));
        static std::once_flag once;
        std::call_once(once, [] { cudaFuncSetAttribute(pare_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDyn); });
        pare_pool_kernel<true><<<N * pare_num_chunks(H, W), 256, kDyn, s>>>(*d);
    } else {
        pare_pool_kernel<false><<<N * pare_num_chunks(H, W), 256, 0, s>>>(*d);
    }
    POCO_LAUNCHED();
    pare_final_kernel<<<N, 256, 0, s>>>(*d);
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_realnvp_run(const poco_realnvp* d, void* stream) {
    POCO_CHECK(d->x && d->params && d->out && d->R > 0, "bad arguments");
    POCO_CHECK(d->D <= RMAXD && d->HID <= RMAXH && d->D + d->CTX <= RMAXIN && (d->CTX == 0 || d->ctx || d->ctx_part), "unsupported flow shape");
    POCO_CHECK(!d->ctx_part || (d->CTX > 0 && d->ctx_group >= 1), "ctx_part needs a conditional flow and ctx_group >= 1");
    realnvp_kernel<<<blocks_for(d->R, RB), 128, 0, static_cast<cudaStream_t>(stream)>>>(*d);
    POCO_LAUNCHED();
    return 0;
}
