// poco_b200 -- SMPL mesh stage on the device: linear blend skinning, the 49 joints of the
// reference wrapper, camera conversions and the 2-D projection of POCO.forward's last step
// (pocolib/models/head/smpl_head.py:12-34, 45-83; smplcam_head.py:34-139; utils/geometry.py:447-463, 480-508).
// The LBS arithmetic is smplx==0.1.28's (requirements.txt:7; lbs.py `lbs` with pose2rot=False) -- a third-party
// package absent from the reference tree, restated from its published algorithm: parity for it is UNPINNED.
//
// Three launches per batch, all fp32, HBM/L2-bound (8.6 MFLOP and 83 KB of vertices per crop):
//   smpl_pose_kernel    one warp per crop: blend coefficients (10 betas + 207 pose features), rest joints from the
//                       host-folded regressor (J = J_template + J_dirs . beta, exact because vertices2joints is
//                       linear in the shaped vertices), the 24-joint rigid chain and the relative transforms A
//   smpl_skin_kernel    one thread per vertex x 8 crops per CTA: v_posed = v_template + dirs^T coef (the blend
//                       directions are read once per 8 crops, coalesced, coefficients broadcast from shared
//                       memory), T = sum_j w_j A_j, v = T [v_posed; 1]
//   smpl_joints_kernel  one CTA per crop: vertex joints, sparse extra regressor (CSR), joint_map, cameras, projection
#include "internal.h"

namespace poco {
namespace {

constexpr int kJ = POCO_SMPL_JOINTS;            // 24
constexpr int kNB = POCO_SMPL_BETAS;            // 10
constexpr int kCoef = kNB + (kJ - 1) * 9;       // 217
constexpr int kCoefPad = 220;
constexpr int kOffA = kCoefPad;                 // [24][12] relative transforms (3x4 row-major)
constexpr int kOffJ = kOffA + kJ * 12;          // [24][3] posed joints
static_assert(kOffJ + kJ * 3 == POCO_SMPL_SCRATCH_FLOATS, "scratch layout");
constexpr int kMaxJointsAll = 64;
constexpr int kCB = 8;                          // crops per CTA in the skinning kernel
constexpr int kKU = 8;                          // blend rows per register-prefetch group
constexpr int kRows = POCO_SMPL_DIR_ROWS;       // 224: the 217 blend rows + zero rows, a whole number of group pairs
constexpr int kGroups = kRows / kKU;
constexpr int kSmplVP = 6912;                   // 6890 vertices of the published model, padded to 128
static_assert(kRows >= kCoef && kRows % (2 * kKU) == 0, "blend table rows");

__global__ void __launch_bounds__(128) smpl_pose_kernel(poco_smpl d) {
    __shared__ float sG[4][kJ][12];
    __shared__ float sJ[4][kJ][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= d.n) return;                       // (whole warps leave; only __syncwarp below)
    const float* R = d.rotmat + (size_t)b * (kJ * 9);
    const float* beta = d.betas + (size_t)b * kNB;
    float* sc = d.scratch + (size_t)b * POCO_SMPL_SCRATCH_FLOATS;
    for (int i = lane; i < kCoefPad; i += 32) {
        float v = 0.f;
        if (i < kNB) {
            v = beta[i];
        } else if (i < kCoef) {                 // pose feature: (R_j - I) of joints 1..23, row-major (lbs.py)
            const int e = i - kNB, rc = e % 9;
            v = R[9 + e] - ((rc == 0 || rc == 4 || rc == 8) ? 1.f : 0.f);
        }
        sc[i] = v;
    }
    const int parent = lane < kJ ? d.model.parents[lane] : -1;
    if (lane < kJ) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = d.model.j_template[lane * 3 + c];
            const float* jd = d.model.j_dirs + (lane * 3 + c) * kNB;
#pragma unroll
            for (int l = 0; l < kNB; ++l) a = fmaf(jd[l], beta[l], a);
            sJ[warp][lane][c] = a;
        }
    }
    __syncwarp();
    if (lane < kJ) {                            // local transform [R_j | J_j - J_parent]
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) sG[warp][lane][r * 4 + c] = R[lane * 9 + r * 3 + c];
            sG[warp][lane][r * 4 + 3] = sJ[warp][lane][r] - (parent >= 0 ? sJ[warp][parent][r] : 0.f);
        }
    }
    __syncwarp();
    for (int i = 1; i < kJ; ++i) {              // G_i = G_parent(i) . L_i  (parents precede children)
        const int p = __shfl_sync(0xffffffffu, parent, i);
        float v = 0.f;
        if (lane < 12) {
            const int r = lane >> 2, c = lane & 3;
            const float* Gp = sG[warp][p];
            const float* Li = sG[warp][i];
            v = Gp[r * 4 + 0] * Li[c] + Gp[r * 4 + 1] * Li[4 + c] + Gp[r * 4 + 2] * Li[8 + c] + (c == 3 ? Gp[r * 4 + 3] : 0.f);
        }
        __syncwarp();
        if (lane < 12) sG[warp][i][lane] = v;
        __syncwarp();
    }
    if (lane < kJ) {                            // A_j = G_j with translation - R_j^G J_j ; posed joint = translation
        const float* G = sG[warp][lane];
        const float j0 = sJ[warp][lane][0], j1 = sJ[warp][lane][1], j2 = sJ[warp][lane][2];
        float* A = sc + kOffA + lane * 12;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            A[r * 4 + 0] = G[r * 4 + 0];
            A[r * 4 + 1] = G[r * 4 + 1];
            A[r * 4 + 2] = G[r * 4 + 2];
            A[r * 4 + 3] = G[r * 4 + 3] - (G[r * 4 + 0] * j0 + G[r * 4 + 1] * j1 + G[r * 4 + 2] * j2);
            sc[kOffJ + lane * 3 + r] = G[r * 4 + 3];
        }
    }
}

// Stages the blend coefficients and relative transforms of a CTA's 8 crops in shared memory.  All global loads of a
// thread are issued before its first shared-memory store (two L2 round trips instead of one per element: the
// element-wise loop spent a third of the kernel here, profiles/r01d_ncu_full_smpl_summary.csv).  Kept out of line so
// its temporaries do not weigh on the register allocation of the blend loop.
static __device__ __noinline__ void smpl_stage_crops(const float* __restrict__ scratch, int b0, int nb,
                                                     float (*sCoef)[kCB], float (*sA)[kJ * 12]) {
    constexpr int kItC = kCB * kRows / 128, kItA = kCB * kJ * 12 / 128;
    static_assert(kCB * kRows % 128 == 0 && kCB * kJ * 12 % 128 == 0, "staging loops are whole");
    {
        float tc[kItC];
#pragma unroll
        for (int it = 0; it < kItC; ++it) {
            const int i = threadIdx.x + it * 128, cb = i / kRows, k = i - cb * kRows;
            tc[it] = (cb < nb && k < kCoef) ? scratch[(size_t)(b0 + cb) * POCO_SMPL_SCRATCH_FLOATS + k] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < kItC; ++it) {
            const int i = threadIdx.x + it * 128, cb = i / kRows, k = i - cb * kRows;
            sCoef[k][cb] = tc[it];
        }
    }
    {
        float ta[kItA];
#pragma unroll
        for (int it = 0; it < kItA; ++it) {
            const int i = threadIdx.x + it * 128, cb = i / (kJ * 12), e = i - cb * (kJ * 12);
            ta[it] = cb < nb ? scratch[(size_t)(b0 + cb) * POCO_SMPL_SCRATCH_FLOATS + kOffA + e] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < kItA; ++it) {
            const int i = threadIdx.x + it * 128, cb = i / (kJ * 12), e = i - cb * (kJ * 12);
            sA[cb][e] = ta[it];
        }
    }
}

// VP != 0: the padded vertex count is a compile-time constant (6912 for the published 6890-vertex model), so the 24
// loads of a group are one base register + immediate offsets; VP == 0: any model, strides from poco_smpl_model.vp.
template <int VP>
__global__ void __launch_bounds__(128) smpl_skin_kernel(poco_smpl d) {
    __shared__ __align__(16) float sCoef[kRows][kCB];
    __shared__ __align__(16) float sA[kCB][kJ * 12];
    const int b0 = blockIdx.y * kCB;
    const int nb = min(kCB, d.n - b0);
    smpl_stage_crops(d.scratch, b0, nb, sCoef, sA);
    __syncthreads();
    const int vp = VP ? VP : d.model.vp;
    const int v = blockIdx.x * 128 + threadIdx.x;       // (vp is a multiple of 128; padded vertices hold zeros)
    float acc[kCB][3];
    {
        const float t0 = d.model.v_template[v], t1 = d.model.v_template[vp + v], t2 = d.model.v_template[2 * vp + v];
#pragma unroll
        for (int cb = 0; cb < kCB; ++cb) { acc[cb][0] = t0; acc[cb][1] = t1; acc[cb][2] = t2; }
    }
    // blend: acc += coef[k] * dirs[k] over the rows of the table (217 + zero rows up to kRows).  The rows come from L2
    // (every CTA of the launch reads the same table), ~700 cycles away: two register buffers of kKU rows alternate,
    // group g+1 is in flight while the FMAs of group g issue, so no load is waited on inside a group.
    const float* dp = d.model.dirs + v;
    const size_t row = VP ? size_t(VP) : size_t(vp);
    float ba[kKU][3], bb[kKU][3];
    auto fetch = [&](float (&buf)[kKU][3], int g) {
        const float* p = dp + (size_t)g * (kKU * 3) * row;
#pragma unroll
        for (int u = 0; u < kKU; ++u) {
#pragma unroll
            for (int c = 0; c < 3; ++c) buf[u][c] = __ldg(p + (u * 3 + c) * row);
        }
    };
    auto blend = [&](const float (&buf)[kKU][3], int g) {
#pragma unroll
        for (int u = 0; u < kKU; ++u) {
            const float4 c0 = *reinterpret_cast<const float4*>(&sCoef[g * kKU + u][0]);
            const float4 c1 = *reinterpret_cast<const float4*>(&sCoef[g * kKU + u][4]);
            const float c[kCB] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int cb = 0; cb < kCB; ++cb) {
                acc[cb][0] = fmaf(c[cb], buf[u][0], acc[cb][0]);
                acc[cb][1] = fmaf(c[cb], buf[u][1], acc[cb][1]);
                acc[cb][2] = fmaf(c[cb], buf[u][2], acc[cb][2]);
            }
        }
    };
    fetch(ba, 0);
#pragma unroll 1
    for (int g = 0; g < kGroups; g += 2) {
        fetch(bb, g + 1);
        blend(ba, g);
        if (g + 2 < kGroups) fetch(ba, g + 2);
        blend(bb, g + 1);
    }
    float w[kJ];
#pragma unroll
    for (int j = 0; j < kJ; ++j) w[j] = __ldg(d.model.weights + (size_t)j * vp + v);
    if (v >= d.model.nv) return;
#pragma unroll
    for (int cb = 0; cb < kCB; ++cb) {
        if (cb >= nb) break;
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
            const float4* a = reinterpret_cast<const float4*>(&sA[cb][j * 12]);
            const float4 a0 = a[0], a1 = a[1], a2 = a[2];
            T[0] = fmaf(w[j], a0.x, T[0]); T[1] = fmaf(w[j], a0.y, T[1]); T[2] = fmaf(w[j], a0.z, T[2]); T[3] = fmaf(w[j], a0.w, T[3]);
            T[4] = fmaf(w[j], a1.x, T[4]); T[5] = fmaf(w[j], a1.y, T[5]); T[6] = fmaf(w[j], a1.z, T[6]); T[7] = fmaf(w[j], a1.w, T[7]);
            T[8] = fmaf(w[j], a2.x, T[8]); T[9] = fmaf(w[j], a2.y, T[9]); T[10] = fmaf(w[j], a2.z, T[10]); T[11] = fmaf(w[j], a2.w, T[11]);
        }
        const float px = acc[cb][0], py = acc[cb][1], pz = acc[cb][2];
        float* o = d.vertices + ((size_t)(b0 + cb) * d.model.nv + v) * 3;
        o[0] = T[0] * px + T[1] * py + T[2] * pz + T[3];
        o[1] = T[4] * px + T[5] * py + T[6] * pz + T[7];
        o[2] = T[8] * px + T[9] * py + T[10] * pz + T[11];
    }
}

__global__ void __launch_bounds__(128) smpl_joints_kernel(poco_smpl d) {
    __shared__ float sAll[kMaxJointsAll][3];
    __shared__ float sCam[8];                   // tx ty tz fx cx cy
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const poco_smpl_model& m = d.model;
    const float* sc = d.scratch + (size_t)b * POCO_SMPL_SCRATCH_FLOATS;
    const float* verts = d.vertices + (size_t)b * m.nv * 3;
    if (tid < kJ * 3) sAll[tid / 3][tid % 3] = sc[kOffJ + tid];
    for (int e = tid; e < m.n_extra_vertex * 3; e += 128) {     // smplx VertexJointSelector
        const int j = e / 3, c = e - j * 3;
        sAll[kJ + j][c] = verts[(size_t)m.extra_vertex_ids[j] * 3 + c];
    }
    for (int r = warp; r < m.n_extra_reg; r += 4) {             // J_regressor_extra (smpl_head.py:24), CSR rows
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int i = m.reg_row_ptr[r] + lane; i < m.reg_row_ptr[r + 1]; i += 32) {
            const float val = m.reg_val[i];
            const float* p = verts + (size_t)m.reg_col[i] * 3;
            s0 = fmaf(val, p[0], s0);
            s1 = fmaf(val, p[1], s1);
            s2 = fmaf(val, p[2], s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            float* o = sAll[kJ + m.n_extra_vertex + r];
            o[0] = s0; o[1] = s1; o[2] = s2;
        }
    }
    if (tid == 0) {
        const float s = d.cam[b * 3 + 0], tx = d.cam[b * 3 + 1], ty = d.cam[b * 3 + 2];
        // convert_weak_perspective_to_perspective with its defaults (geometry.py:447-463)
        const float ctz = 2.f * 5000.f / (224.f * s + 1e-9f);
        if (d.cam_t) { d.cam_t[b * 3 + 0] = tx; d.cam_t[b * 3 + 1] = ty; d.cam_t[b * 3 + 2] = ctz; }
        if (d.cliff) {                          // convert_pare_to_full_img_cam (smplcam_head.py:123-139)
            const float f = d.focal_length[b], bh = d.bbox_scale[b] * 200.f, iw = d.img_w[b], ih = d.img_h[b];
            const float r = bh / 224.f;
            const float tz = 2.f * f / (r * 224.f * s);
            const float cx = 2.f * (d.bbox_center[b * 2 + 0] - iw / 2.f) / (s * bh);
            const float cy = 2.f * (d.bbox_center[b * 2 + 1] - ih / 2.f) / (s * bh);
            sCam[0] = tx + cx; sCam[1] = ty + cy; sCam[2] = tz;
            sCam[3] = f; sCam[4] = iw / 2.f; sCam[5] = ih / 2.f;
            if (d.fullimg_cam_t) {
                d.fullimg_cam_t[b * 3 + 0] = sCam[0]; d.fullimg_cam_t[b * 3 + 1] = sCam[1]; d.fullimg_cam_t[b * 3 + 2] = sCam[2];
            }
        } else {
            sCam[0] = tx; sCam[1] = ty; sCam[2] = ctz;
            sCam[3] = d.focal_default; sCam[4] = 0.f; sCam[5] = 0.f;
        }
    }
    __syncthreads();
    const int n_all = kJ + m.n_extra_vertex + m.n_extra_reg;
    for (int i = tid; i < m.n_joints_out; i += 128) {
        const int j = min(max(m.joint_map[i], 0), n_all - 1);
        const float x = sAll[j][0], y = sAll[j][1], z = sAll[j][2];
        float* o3 = d.joints3d + ((size_t)b * m.n_joints_out + i) * 3;
        o3[0] = x; o3[1] = y; o3[2] = z;
        if (d.joints2d) {                       // perspective_projection, identity rotation
            const float qx = x + sCam[0], qy = y + sCam[1], qz = z + sCam[2];
            const float px = qx / qz, py = qy / qz, pz = qz / qz;
            float u = sCam[3] * px + sCam[4] * pz, w = sCam[3] * py + sCam[5] * pz;
            if (d.normalize_joints2d) { u = u / (d.img_res / 2.f); w = w / (d.img_res / 2.f); }
            d.joints2d[((size_t)b * m.n_joints_out + i) * 2 + 0] = u;
            d.joints2d[((size_t)b * m.n_joints_out + i) * 2 + 1] = w;
        }
    }
}

}  // namespace
}  // namespace poco

using namespace poco;

extern "C" int poco_smpl_run(const poco_smpl* d, void* stream) {
    const poco_smpl_model& m = d->model;
    POCO_CHECK(d->n > 0, "empty batch");
    POCO_CHECK(d->rotmat && d->betas && d->cam && d->scratch && d->vertices && d->joints3d, "null pointer");
    POCO_CHECK(m.v_template && m.dirs && m.weights && m.j_template && m.j_dirs && m.parents && m.joint_map, "null model pointer");
    POCO_CHECK(m.nv > 0 && m.vp >= m.nv && m.vp % 128 == 0, "model: vp must be nv rounded up to a multiple of 128");
    POCO_CHECK(m.n_extra_vertex >= 0 && m.n_extra_reg >= 0 && kJ + m.n_extra_vertex + m.n_extra_reg <= kMaxJointsAll,
               "model: too many joints");
    POCO_CHECK(m.n_extra_vertex == 0 || m.extra_vertex_ids, "model: null extra_vertex_ids");
    POCO_CHECK(m.n_extra_reg == 0 || (m.reg_row_ptr && m.reg_col && m.reg_val), "model: null extra regressor");
    POCO_CHECK(m.n_joints_out > 0, "model: no output joints");
    POCO_CHECK(!d->cliff || (d->focal_length && d->bbox_scale && d->bbox_center && d->img_w && d->img_h),
               "cliff: null camera input");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    smpl_pose_kernel<<<(d->n + 3) / 4, 128, 0, s>>>(*d);
    POCO_LAUNCHED();
    const dim3 grid(m.vp / 128, (d->n + kCB - 1) / kCB);
    if (m.vp == kSmplVP)
        smpl_skin_kernel<kSmplVP><<<grid, 128, 0, s>>>(*d);
    else
        smpl_skin_kernel<0><<<grid, 128, 0, s>>>(*d);
    POCO_LAUNCHED();
    smpl_joints_kernel<<<d->n, 128, 0, s>>>(*d);
    POCO_LAUNCHED();
    return 0;
}
