// poco_b200 -- bandwidth-bound layout / pooling / resampling kernels on the planar-8 padded fp16
// activation layout, plus the CUDA-core debug convolution.  All of them move 16 bytes (8 channels
// of one pixel) per thread access; consecutive threads touch consecutive pixels of a plane, so a
// warp reads / writes 512 contiguous bytes.
#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace poco {

namespace {

struct Act {
    __half* data;
    long long plane;
    int C, N, H, W;
    __half* lo;         // split-precision mode: rounding residual tensor (value = data + lo), else nullptr
};
inline Act mk(const poco_act& a) { return Act{static_cast<__half*>(a.data), a.plane_stride, a.C, a.N, a.H, a.W, static_cast<__half*>(a.lo)}; }

__device__ __forceinline__ long long pix_index(const Act& a, int n, int y, int x) {
    return (long long)n * (a.H + 2) * (a.W + 2) + (long long)(y + 1) * (a.W + 2) + (x + 1);
}
__device__ __forceinline__ uint4 ld16(const Act& a, int plane, long long pix) {
    return *reinterpret_cast<const uint4*>(a.data + ((long long)plane * a.plane + pix) * 8);
}
__device__ __forceinline__ void st16(const Act& a, int plane, long long pix, uint4 v) {
    *reinterpret_cast<uint4*>(a.data + ((long long)plane * a.plane + pix) * 8) = v;
}
__device__ __forceinline__ void unpack8(uint4 v, float* f) {
    float2 t;
    t = unpack_half2(v.x); f[0] = t.x; f[1] = t.y;
    t = unpack_half2(v.y); f[2] = t.x; f[3] = t.y;
    t = unpack_half2(v.z); f[4] = t.x; f[5] = t.y;
    t = unpack_half2(v.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 v;
    v.x = pack_half2(f[0], f[1]); v.y = pack_half2(f[2], f[3]);
    v.z = pack_half2(f[4], f[5]); v.w = pack_half2(f[6], f[7]);
    return v;
}
// value of 8 channels of one pixel: hi (+ lo in split-precision mode; the branch is uniform per launch)
__device__ __forceinline__ void ld8f(const Act& a, int plane, long long pix, float* f) {
    unpack8(ld16(a, plane, pix), f);
    if (a.lo != nullptr) {
        float l[8];
        unpack8(*reinterpret_cast<const uint4*>(a.lo + ((long long)plane * a.plane + pix) * 8), l);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += l[i];
    }
}
// store fp32 values as hi = fp16(v) (+ lo = fp16(v - hi))
__device__ __forceinline__ void st8f(const Act& a, int plane, long long pix, const float* f) {
    const uint4 h = pack8(f);
    st16(a, plane, pix, h);
    if (a.lo != nullptr) {
        float hf[8], l[8];
        unpack8(h, hf);
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = f[i] - hf[i];
        *reinterpret_cast<uint4*>(a.lo + ((long long)plane * a.plane + pix) * 8) = pack8(l);
    }
}

// ------------------------------------------------------------------------------------------------
// debug convolution (CUDA cores, fp32 accumulate): one thread = one output pixel x 8 output channels
// ------------------------------------------------------------------------------------------------
__global__ void conv_ref_kernel(Act in, Act out, const __half* __restrict__ w, const float* __restrict__ bias,
                                const __half* __restrict__ res, const __half* __restrict__ res_lo, long long res_plane,
                                int kh, int kw, int stride, int pad, int relu) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)out.N * out.H * out.W;
    if (idx >= npix) return;
    const int co8 = blockIdx.y;
    const int x = int(idx % out.W), y = int((idx / out.W) % out.H), n = int(idx / ((long long)out.W * out.H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias[co8 * 8 + i];
    const int cin8 = in.C >> 3;
    for (int t = 0; t < kh * kw; ++t) {
        const int yi = y * stride + t / kw - pad, xi = x * stride + t % kw - pad;
        if (yi < 0 || yi >= in.H || xi < 0 || xi >= in.W) continue;
        const long long ip = pix_index(in, n, yi, xi);
        for (int c8 = 0; c8 < cin8; ++c8) {
            float a[8];
            ld8f(in, c8, ip, a);
            const uint4* wp = reinterpret_cast<const uint4*>(w + ((long long)(t * cin8 + c8) * out.C + co8 * 8) * 8);
            const uint4* wl = wp + (long long)kh * kw * cin8 * out.C;      // split mode: W_lo follows W_hi
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float wf[8];
                unpack8(__ldg(wp + o), wf);
                if (in.lo != nullptr) {
                    float wlo[8];
                    unpack8(__ldg(wl + o), wlo);
#pragma unroll
                    for (int i = 0; i < 8; ++i) wf[i] += wlo[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[o] = fmaf(a[i], wf[i], acc[o]);
            }
        }
    }
    const long long op = pix_index(out, n, y, x);
    if (relu == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    if (res != nullptr) {
        float r[8];
        unpack8(*reinterpret_cast<const uint4*>(res + ((long long)co8 * res_plane + op) * 8), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += r[i];
        if (res_lo != nullptr) {
            unpack8(*reinterpret_cast<const uint4*>(res_lo + ((long long)co8 * res_plane + op) * 8), r);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += r[i];
        }
    }
    if (relu == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    st8f(out, co8, op, acc);
}

// ------------------------------------------------------------------------------------------------
// Thread mapping of the element-wise kernels: a 256-thread block covers 256 / slots image rows, where
// slots = W rounded up to a power of two; thread t handles column t % slots of row t / slots.  All index
// math is 32-bit with one division per thread (the first version's 64-bit div / mod per pixel kept
// these kernels at ~1 TB/s).  blockIdx.y = plane group.
struct RowMap {
    int slots_log2;
    int rows;           // N * H
};
inline RowMap row_map(int N, int H, int W) {
    int l = 0;
    while ((1 << l) < W) ++l;
    return RowMap{std::min(l, 8), N * H};
}
inline unsigned row_blocks(const RowMap& m) { return unsigned((m.rows + (256 >> m.slots_log2) - 1) / (256 >> m.slots_log2)); }
__device__ __forceinline__ bool row_coords(const RowMap& m, int H, int W, int& n, int& y, int& x) {
    x = threadIdx.x & ((1 << m.slots_log2) - 1);
    const int row = blockIdx.x * (256 >> m.slots_log2) + (threadIdx.x >> m.slots_log2);
    if (row >= m.rows || x >= W) return false;
    n = row / H;
    y = row - n * H;
    return true;
}

__global__ void pack_image_kernel(const float* __restrict__ img, Act out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)out.N * out.H * out.W;
    if (idx >= npix) return;
    const int x = int(idx % out.W), y = int((idx / out.W) % out.H), n = int(idx / ((long long)out.W * out.H));
    const long long hw = (long long)out.H * out.W;
    const float* src = img + (long long)n * 3 * hw + (long long)y * out.W + x;
    float f[8] = {src[0], src[hw], src[2 * hw], 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long op = pix_index(out, n, y, x);
    st8f(out, 0, op, f);
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int pl = 1; pl < (out.C >> 3); ++pl) st8f(out, pl, op, z);
}

// im2col of the 3x3 / stride-2 / pad-1 stem window: 27 taps x channels -> 32 fp16 channels per output pixel
__global__ void __launch_bounds__(256) pack_stem_kernel(const float* __restrict__ img, Act out, RowMap m) {
    int n, y, x;
    if (!row_coords(m, out.H, out.W, n, y, x)) return;
    const int Hi = 2 * out.H, Wi = 2 * out.W;
    const long long hw = (long long)Hi * Wi;
    const float* src = img + (long long)n * 3 * hw;
    float f[32];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int yi = 2 * y + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int xi = 2 * x + s - 1;
            const bool ok = yi >= 0 && yi < Hi && xi >= 0 && xi < Wi;
            const long long o = (long long)yi * Wi + xi;
#pragma unroll
            for (int c = 0; c < 3; ++c) f[(r * 3 + s) * 3 + c] = ok ? __ldg(src + c * hw + o) : 0.f;
        }
    }
#pragma unroll
    for (int k = 27; k < 32; ++k) f[k] = 0.f;
    const long long op = pix_index(out, n, y, x);
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) st8f(out, pl, op, f + pl * 8);
}

struct FuseArgs {
    Act out;
    Act in[POCO_MAX_FUSE_INPUTS];
    int shift[POCO_MAX_FUSE_INPUTS];
    int n_in, relu;
};
// Occupancy is what these two HBM-bound kernels live on (round 2 ncu: 74 / 64 registers per thread held them at 3-4
// blocks per SM, 34-44 % of the warp slots, and at 40 % of the HBM rate): the precision mode is a template parameter so
// that the fp16 instantiation does not carry the lo-tensor registers, one plane is in flight per loop iteration, and
// __launch_bounds__ keeps 5-6 blocks (62-75 % of the slots) resident in fp16 mode.
template <bool SPLIT>
__global__ void __launch_bounds__(256, SPLIT ? 4 : 5) fuse_sum_kernel(FuseArgs a, RowMap m, int planes_per_thread) {
    int n, y, x;
    if (!row_coords(m, a.out.H, a.out.W, n, y, x)) return;
    int pin[POCO_MAX_FUSE_INPUTS];          // pixel index inside a plane (< 2^31: checked by the launcher)
#pragma unroll
    for (int k = 0; k < POCO_MAX_FUSE_INPUTS; ++k)
        pin[k] = k < a.n_in ? int(pix_index(a.in[k], n, y >> a.shift[k], x >> a.shift[k])) : 0;
    const long long po = pix_index(a.out, n, y, x);
    const int pl0 = blockIdx.y * planes_per_thread;
    for (int pl = pl0; pl < pl0 + planes_per_thread; ++pl) {
        uint4 v[POCO_MAX_FUSE_INPUTS];
#pragma unroll
        for (int k = 0; k < POCO_MAX_FUSE_INPUTS; ++k)
            if (k < a.n_in) v[k] = ld16(a.in[k], pl, pin[k]);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < POCO_MAX_FUSE_INPUTS; ++k)
            if (k < a.n_in) {
                float f[8];
                unpack8(v[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += f[i];
            }
        if (SPLIT) {
#pragma unroll
            for (int k = 0; k < POCO_MAX_FUSE_INPUTS; ++k)
                if (k < a.n_in) v[k] = *reinterpret_cast<const uint4*>(a.in[k].lo + ((long long)pl * a.in[k].plane + pin[k]) * 8);
#pragma unroll
            for (int k = 0; k < POCO_MAX_FUSE_INPUTS; ++k)
                if (k < a.n_in) {
                    float f[8];
                    unpack8(v[k], f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += f[i];
                }
        }
        if (a.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaxf(acc[i], 0.f);
        }
        st8f(a.out, pl, po, acc);
    }
}

// bilinear x2, align_corners=True: src = dst * (in-1)/(out-1)  (matches aten upsample_bilinear2d)
template <bool SPLIT>
__global__ void __launch_bounds__(256, SPLIT ? 4 : 6) upsample2x_kernel(Act in, Act out, float sy, float sx, RowMap m, int planes_per_thread) {
    int n, y, x;
    if (!row_coords(m, out.H, out.W, n, y, x)) return;
    const float fy = sy * y, fx = sx * x;
    const int y0 = min(int(fy), in.H - 1), x0 = min(int(fx), in.W - 1);
    const int y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const int p00 = int(pix_index(in, n, y0, x0)), p01 = int(pix_index(in, n, y0, x1));     // (< 2^31: checked by the launcher)
    const int p10 = int(pix_index(in, n, y1, x0)), p11 = int(pix_index(in, n, y1, x1));
    const long long po = pix_index(out, n, y, x);
    const int pl0 = blockIdx.y * planes_per_thread;
    for (int pl = pl0; pl < pl0 + planes_per_thread; ++pl) {
        float a[8], b[8], c[8], d[8], o[8];
        if (SPLIT) {            // split-precision mode: interpolate hi + lo in fp32
            ld8f(in, pl, p00, a); ld8f(in, pl, p01, b); ld8f(in, pl, p10, c); ld8f(in, pl, p11, d);
        } else {
            const uint4 v0 = ld16(in, pl, p00), v1 = ld16(in, pl, p01), v2 = ld16(in, pl, p10), v3 = ld16(in, pl, p11);
            unpack8(v0, a); unpack8(v1, b); unpack8(v2, c); unpack8(v3, d);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = hy * (hx * a[i] + lx * b[i]) + ly * (hx * c[i] + lx * d[i]);
        st8f(out, pl, po, o);
    }
}

__global__ void maxpool_kernel(Act in, Act out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)out.N * out.H * out.W;
    if (idx >= npix) return;
    const int pl = blockIdx.y;
    const int x = int(idx % out.W), y = int((idx / out.W) % out.H), n = int(idx / ((long long)out.W * out.H));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yi = 2 * y + dy, xi = 2 * x + dx;
            if (yi < 0 || yi >= in.H || xi < 0 || xi >= in.W) continue;
            float f[8];
            ld8f(in, pl, pix_index(in, n, yi, xi), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
        }
    st8f(out, pl, pix_index(out, n, y, x), m);
}

// one warp per (crop, plane): mean over H*W of 8 channels
__global__ void avgpool_kernel(Act in, float* __restrict__ out, long long ld) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int planes = in.C >> 3;
    if (warp >= in.N * planes) return;
    const int n = warp / planes, pl = warp % planes;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int hw = in.H * in.W;
    for (int i = lane; i < hw; i += 32) {
        float f[8];
        ld8f(in, pl, pix_index(in, n, i / in.W, i % in.W), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[k] = v;
    }
    if (lane < 8) out[(long long)n * ld + pl * 8 + lane] = acc[lane] / float(hw);
}

__global__ void unpack_kernel(Act in, float* __restrict__ out, int c_valid) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)in.N * in.H * in.W;
    if (idx >= npix) return;
    const int pl = blockIdx.y;
    const int x = int(idx % in.W), y = int((idx / in.W) % in.H), n = int(idx / ((long long)in.W * in.H));
    float f[8];
    ld8f(in, pl, pix_index(in, n, y, x), f);
    const long long hw = (long long)in.H * in.W;
    for (int i = 0; i < 8; ++i) {
        const int c = pl * 8 + i;
        if (c < c_valid) out[((long long)n * c_valid + c) * hw + (long long)y * in.W + x] = f[i];
    }
}

inline unsigned blocks_for(long long n, int bs) { return unsigned((n + bs - 1) / bs); }

}  // namespace

int check_act(const poco_act& a, const char* what) {
    POCO_CHECK(a.data != nullptr, std::string(what) + ": null data");
    POCO_CHECK(a.C > 0 && a.C % 8 == 0, std::string(what) + ": channels must be a positive multiple of 8");
    POCO_CHECK(a.N > 0 && a.H > 0 && a.W > 0, std::string(what) + ": empty tensor");
    POCO_CHECK(a.plane_stride >= int64_t(a.N) * (a.H + 2) * (a.W + 2), std::string(what) + ": plane stride too small");
    POCO_CHECK((reinterpret_cast<uintptr_t>(a.data) & 15) == 0, std::string(what) + ": data must be 16-byte aligned");
    POCO_CHECK((reinterpret_cast<uintptr_t>(a.lo) & 15) == 0, std::string(what) + ": lo must be 16-byte aligned");
    return 0;
}

int conv_ref_launch(const poco_conv* d, cudaStream_t s) {
    POCO_CHECK(d->in.N == d->out.N, "batch mismatch");
    POCO_CHECK((d->in.lo != nullptr) == (d->out.lo != nullptr) && (!d->in.lo || !d->residual || d->residual_lo),
               "split precision: in.lo, out.lo and residual_lo must be given together");
    POCO_CHECK((d->in.H + 2 * d->pad - d->kh) / d->stride + 1 == d->out.H &&
                   (d->in.W + 2 * d->pad - d->kw) / d->stride + 1 == d->out.W,
               "output geometry does not match the convolution");
    const long long npix = (long long)d->out.N * d->out.H * d->out.W;
    dim3 grid(blocks_for(npix, 128), d->out.C / 8);
    conv_ref_kernel<<<grid, 128, 0, s>>>(mk(d->in), mk(d->out), static_cast<const __half*>(d->weight), d->bias,
                                        static_cast<const __half*>(d->residual), static_cast<const __half*>(d->residual_lo),
                                        d->res_plane_stride, d->kh, d->kw, d->stride, d->pad, d->relu);
    POCO_LAUNCHED();
    return 0;
}

}  // namespace poco

using namespace poco;

extern "C" int poco_pack_image_run(const poco_pack_image* d, void* stream) {
    if (check_act(d->out, "out")) return 1;
    POCO_CHECK(d->img != nullptr, "null image");
    if (d->im2col) {
        POCO_CHECK(d->out.C == 32 && d->out.W <= 256, "im2col stem packing writes 32 channels, W <= 256");
        const RowMap m = row_map(d->out.N, d->out.H, d->out.W);
        pack_stem_kernel<<<row_blocks(m), 256, 0, static_cast<cudaStream_t>(stream)>>>(d->img, mk(d->out), m);
        POCO_LAUNCHED();
        return 0;
    }
    const long long npix = (long long)d->out.N * d->out.H * d->out.W;
    pack_image_kernel<<<blocks_for(npix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d->img, mk(d->out));
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_fuse_sum_run(const poco_fuse_sum* d, void* stream) {
    if (check_act(d->out, "out")) return 1;
    POCO_CHECK(d->n_in >= 1 && d->n_in <= POCO_MAX_FUSE_INPUTS, "n_in out of range");
    FuseArgs a;
    a.out = mk(d->out);
    a.n_in = d->n_in;
    a.relu = d->relu;
    POCO_CHECK(d->out.W <= 256, "fuse_sum: W <= 256");
    for (int k = 0; k < d->n_in; ++k) {
        if (check_act(d->in[k], "in")) return 1;
        POCO_CHECK(d->in[k].C == d->out.C && d->in[k].N == d->out.N, "fuse input channel/batch mismatch");
        POCO_CHECK((d->in[k].lo != nullptr) == (d->out.lo != nullptr), "fuse_sum: inputs and output must share one precision mode");
        POCO_CHECK(d->shift[k] >= 0 && (d->in[k].H << d->shift[k]) == d->out.H && (d->in[k].W << d->shift[k]) == d->out.W,
                   "fuse input resolution mismatch");
        a.in[k] = mk(d->in[k]);
        a.shift[k] = d->shift[k];
    }
    const RowMap m = row_map(d->out.N, d->out.H, d->out.W);
    const int planes = d->out.C / 8;
    const int ppt = planes % 4 == 0 ? 4 : (planes % 2 == 0 ? 2 : 1);      // planes per thread: independent loads in flight
    dim3 grid(row_blocks(m), planes / ppt);
    POCO_CHECK(d->out.plane_stride < (1ll << 31), "fuse_sum: plane too large for 32-bit pixel indices");
    if (d->out.lo != nullptr) fuse_sum_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, m, ppt);
    else fuse_sum_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, m, ppt);
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_upsample2x_run(const poco_upsample2x* d, void* stream) {
    if (check_act(d->in, "in") || check_act(d->out, "out")) return 1;
    POCO_CHECK(d->out.H == 2 * d->in.H && d->out.W == 2 * d->in.W && d->out.C == d->in.C && d->out.N == d->in.N,
               "upsample geometry mismatch");
    const float sy = d->out.H > 1 ? float(d->in.H - 1) / float(d->out.H - 1) : 0.f;
    const float sx = d->out.W > 1 ? float(d->in.W - 1) / float(d->out.W - 1) : 0.f;
    POCO_CHECK(d->out.C % 16 == 0 && d->out.W <= 256, "upsample2x: channels must be a multiple of 16 and W <= 256");
    POCO_CHECK((d->in.lo != nullptr) == (d->out.lo != nullptr), "upsample2x: input and output must share one precision mode");
    const RowMap m = row_map(d->out.N, d->out.H, d->out.W);
    const int planes = d->out.C / 8;
    const int ppt = planes % 4 == 0 ? 4 : 2;
    dim3 grid(row_blocks(m), planes / ppt);
    POCO_CHECK(d->out.plane_stride < (1ll << 31), "upsample2x: plane too large for 32-bit pixel indices");
    if (d->out.lo != nullptr) upsample2x_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mk(d->in), mk(d->out), sy, sx, m, ppt);
    else upsample2x_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mk(d->in), mk(d->out), sy, sx, m, ppt);
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_maxpool_run(const poco_maxpool* d, void* stream) {
    if (check_act(d->in, "in") || check_act(d->out, "out")) return 1;
    POCO_CHECK(d->out.H == (d->in.H + 2 - 3) / 2 + 1 && d->out.W == (d->in.W + 2 - 3) / 2 + 1 && d->out.C == d->in.C,
               "maxpool geometry mismatch");
    POCO_CHECK((d->in.lo != nullptr) == (d->out.lo != nullptr), "maxpool: input and output must share one precision mode");
    const long long npix = (long long)d->out.N * d->out.H * d->out.W;
    dim3 grid(blocks_for(npix, 256), d->out.C / 8);
    maxpool_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mk(d->in), mk(d->out));
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_avgpool_run(const poco_avgpool* d, void* stream) {
    if (check_act(d->in, "in")) return 1;
    POCO_CHECK(d->out != nullptr && d->ld >= d->in.C, "bad output");
    const long long warps = (long long)d->in.N * (d->in.C / 8);
    avgpool_kernel<<<blocks_for(warps * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mk(d->in), d->out, d->ld);
    POCO_LAUNCHED();
    return 0;
}

extern "C" int poco_unpack_run(const poco_unpack* d, void* stream) {
    if (check_act(d->in, "in")) return 1;
    POCO_CHECK(d->out != nullptr && d->c_valid > 0 && d->c_valid <= d->in.C, "bad output");
    const long long npix = (long long)d->in.N * d->in.H * d->in.W;
    dim3 grid(blocks_for(npix, 256), (d->c_valid + 7) / 8);
    unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mk(d->in), d->out, d->c_valid);
    POCO_LAUNCHED();
    return 0;
}
