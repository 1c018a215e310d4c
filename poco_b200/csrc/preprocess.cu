// poco_b200 -- the step right before the hot path: per-detection crop + normalisation on the GPU.
//
// Replaces, for a frame that is already in device memory, the per-detection CPU loop of the reference
// (pocolib/core/tester.py:181-212, demo_utils.py:60-80): get_single_image_crop_demo
// (utils/vibe_image_utils.py:233-267 = gen_trans_from_patch_cv :58-91 + cv2.warpAffine INTER_LINEAR /
// BORDER_CONSTANT :104-105 + ToTensor / Normalize :343-352) and calculate_bbox_info / calculate_focal_length
// (utils/image_utils.py:171-187).
//
// cv2.warpAffine on uint8 is integer arithmetic and is reproduced bit for bit: source coordinates in 1/1024 pixel
// (rint of float64 products), rounded to 1/32 pixel, bilinear weights (32-fx)(32-fy)*32 ... summing to 32768,
// result (sum + 16384) >> 15, taps outside the frame read 0.  One thread = one output pixel (3 channels); the
// frame is read through L1/L2 (neighbouring output pixels share source taps), the f32 NCHW result is written
// with fully coalesced stores.  HBM-bound: algorithmic bytes = N * 3 * crop^2 * 4 written (+ the touched frame
// region read once).
#include "common.cuh"
#include "internal.h"

namespace poco {

namespace {

struct CropArgs {
    const uint8_t* frame;
    int H, W;
    const float* boxes;
    int n, crop;
    double scale;
    float* img;
    float* bbox_info;
    float* focal_length;
    float* scale_out;
    float* center;
    float* orig_shape;
};

// inverse affine of one detection: follows gen_trans_from_patch_cv (float32 control points, float64 solve) and the
// inversion inside cv::warpAffine; m[0..5] = row-major 2x3
__device__ void inverse_affine(double cx, double cy, double bw, double bh, double scale, int crop, double* m) {
    const double src_w = bw * scale, src_h = bh * scale;
    const float p0x = float(cx), p0y = float(cy);
    const float p1y = float(cy + double(float(src_h * 0.5)));
    const float p2x = float(cx + double(float(src_w * 0.5)));
    const double half = double(float(crop * 0.5));
    // (explicit _rn intrinsics: no FMA contraction, the products and sums round exactly like the host code of cv2 / numpy)
    const double a = half / (double(p2x) - double(p0x));
    const double d = half / (double(p1y) - double(p0y));
    double f[6] = {a, 0.0, __dsub_rn(half, __dmul_rn(a, double(p0x))), 0.0, d, __dsub_rn(half, __dmul_rn(d, double(p0y)))};
    double D = __dsub_rn(__dmul_rn(f[0], f[4]), __dmul_rn(f[1], f[3]));
    D = D != 0.0 ? 1.0 / D : 0.0;
    const double A11 = __dmul_rn(f[4], D), A22 = __dmul_rn(f[0], D);
    f[0] = A11;
    f[1] = __dmul_rn(f[1], -D);
    f[3] = __dmul_rn(f[3], -D);
    f[4] = A22;
    const double b1 = __dsub_rn(__dmul_rn(-f[0], f[2]), __dmul_rn(f[1], f[5]));
    const double b2 = __dsub_rn(__dmul_rn(-f[3], f[2]), __dmul_rn(f[4], f[5]));
    f[2] = b1;
    f[5] = b2;
#pragma unroll
    for (int i = 0; i < 6; ++i) m[i] = f[i];
}

__global__ void __launch_bounds__(256) crop_kernel(CropArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    __shared__ double m[6];             // one affine per block (its three fp64 divisions are the expensive part)
    if (threadIdx.x == 0) {
        const float* b = a.boxes + 4 * n;
        inverse_affine(double(b[0]), double(b[1]), double(b[2]), double(b[3]), a.scale, a.crop, m);
    }
    __syncthreads();
    if (x >= a.crop) return;
    // cv::warpAffine: X0 / Y0 per destination row, adelta / bdelta per destination column, AB_SCALE = 1024,
    // round_delta = 16, then >> (AB_BITS - INTER_BITS); __double2int_rn = cvRound (round half to even)
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m[0], double(x)), 1024.0));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m[3], double(x)), 1024.0));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], double(y)), m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], double(y)), m[5]), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const bool x0ok = sx >= 0 && sx < a.W, x1ok = sx + 1 >= 0 && sx + 1 < a.W;
    const bool y0ok = sy >= 0 && sy < a.H, y1ok = sy + 1 >= 0 && sy + 1 < a.H;
    const uint8_t* r0 = a.frame + (long long)sy * a.W * 3;
    const uint8_t* r1 = r0 + (long long)a.W * 3;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    const long long plane = (long long)a.crop * a.crop;
    float* out = a.img + (long long)n * 3 * plane + (long long)y * a.crop + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int p00 = (y0ok && x0ok) ? r0[sx * 3 + c] : 0, p01 = (y0ok && x1ok) ? r0[(sx + 1) * 3 + c] : 0;
        const int p10 = (y1ok && x0ok) ? r1[sx * 3 + c] : 0, p11 = (y1ok && x1ok) ? r1[(sx + 1) * 3 + c] : 0;
        const int v = (w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 16384) >> 15;
        // ToTensor (/255) then Normalize ((t - mean) / std), both IEEE fp32 like torch on the CPU
        out[c * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(float(v), 255.0f), mean[c]), stdv[c]);
    }
}

// per-detection scalars of the batch dict (tester.py:195-212), float64 arithmetic like the numpy reference
__global__ void crop_meta_kernel(CropArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.n) return;
    const float* b = a.boxes + 4 * n;
    const double cx = b[0], cy = b[1], bw = b[2], bh = b[3];
    const double s = fmax(bw, bh) / 200.0;
    const double f = sqrt(double(a.W) * double(a.W) + double(a.H) * double(a.H));
    if (a.bbox_info) {
        const double bb = s * 200.0;
        a.bbox_info[3 * n + 0] = float((cx - a.W / 2.0) / f * 2.8);
        a.bbox_info[3 * n + 1] = float((cy - a.H / 2.0) / f * 2.8);
        a.bbox_info[3 * n + 2] = float((bb - 0.24 * f) / (0.06 * f));
    }
    if (a.focal_length) a.focal_length[n] = float(f);
    if (a.scale_out) a.scale_out[n] = float(s);
    if (a.center) { a.center[2 * n] = float(cx); a.center[2 * n + 1] = float(cy); }
    if (a.orig_shape) { a.orig_shape[2 * n] = float(a.H); a.orig_shape[2 * n + 1] = float(a.W); }
}

}  // namespace

}  // namespace poco

using namespace poco;

extern "C" int poco_crop_run(const poco_crop* d, void* stream) {
    POCO_CHECK(d->frame && d->boxes && d->img, "null pointer");
    POCO_CHECK(d->frame_h > 0 && d->frame_w > 0 && d->n > 0 && d->crop > 0 && d->crop <= 1024, "bad geometry");
    POCO_CHECK(d->scale > 0.f, "bbox scale must be positive");
    POCO_CHECK((long long)d->frame_h * d->frame_w < (1ll << 29), "frame too large for 32-bit fixed-point coordinates");
    CropArgs a{d->frame, d->frame_h, d->frame_w, d->boxes, d->n, d->crop, double(d->scale), d->img,
               d->bbox_info, d->focal_length, d->scale_out, d->center, d->orig_shape};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid((d->crop + 255) / 256, d->crop, d->n);
    crop_kernel<<<grid, 256, 0, s>>>(a);
    POCO_LAUNCHED();
    crop_meta_kernel<<<(d->n + 127) / 128, 128, 0, s>>>(a);
    POCO_LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Uncertainty post-processing, the step right after the hot path: POCOUtils.prepare_uncert
// (pocolib/utils/poco_utils.py:63-94; optional get_kinematic_uncert :21-25 and 1 - var) followed by
// get_global_uncert (:50-61) as pocolib/core/tester.py:243-245 / :418-421 call them -- on the device, so the
// per-batch device -> host synchronisation of the reference loop disappears.  One thread per crop (24 floats).
// ------------------------------------------------------------------------------------------------------------
namespace poco {
namespace {

__constant__ int kSmplParent[24] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};

__global__ void uncert_post_kernel(poco_uncert_post d) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= d.n) return;
    float v[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i] = d.var[(long long)n * 24 + i];
    if (d.kinematic) {                  // children accumulate the (already updated) value of their parent
#pragma unroll
        for (int i = 1; i < 24; ++i) v[i] = __fadd_rn(v[i], v[kSmplParent[i]]);
    }
    if (d.return_conf) {
#pragma unroll
        for (int i = 0; i < 24; ++i) v[i] = __fsub_rn(1.f, v[i]);
    }
    if (d.prepared) {
#pragma unroll
        for (int i = 0; i < 24; ++i) d.prepared[(long long)n * 24 + i] = v[i];
    }
    // get_global_uncert works on a copy: rows whose first entry exceeds the threshold become all ones
    const float thr = d.cliff ? 2.f * d.sensitivity_threshold : d.sensitivity_threshold;
    if (v[0] > thr) {
#pragma unroll
        for (int i = 0; i < 24; ++i) v[i] = 1.f;
    }
    if (d.thresholded) {
#pragma unroll
        for (int i = 0; i < 24; ++i) d.thresholded[(long long)n * 24 + i] = v[i];
    }
    if (d.global_var) {
        if (d.cliff) {
            d.global_var[n] = v[0];
        } else {                        // numpy float32 mean of 24 values: pairwise sum in 8 lanes, then / 24
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = v[i];
#pragma unroll
            for (int i = 8; i < 24; ++i) r[i & 7] = __fadd_rn(r[i & 7], v[i]);
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
            d.global_var[n] = __fdiv_rn(s, 24.f);
        }
    }
}

}  // namespace
}  // namespace poco

extern "C" int poco_uncert_post_run(const poco_uncert_post* d, void* stream) {
    POCO_CHECK(d->var != nullptr && d->n > 0, "bad arguments");
    POCO_CHECK(d->prepared || d->thresholded || d->global_var, "no output requested");
    poco::uncert_post_kernel<<<(d->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*d);
    POCO_LAUNCHED();
    return 0;
}
