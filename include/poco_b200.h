/* poco_b200 -- C ABI of the B200-native POCO per-crop inference hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own -- every dense
 * op on the path is a torch.nn call (pocolib/models/poco.py:99-129) -- so each entry point below
 * names the reference computation it replaces.  All pointers are DEVICE pointers unless stated,
 * `stream` is a cudaStream_t passed as void*, every call is asynchronous on that stream, returns
 * 0 on success and a non-zero code otherwise (never throws); poco_last_error() gives the message
 * of the last failure on the calling thread.  No torch types cross this boundary.
 *
 * Activation tensors use the "planar-8 padded" fp16 layout  [C/8][N][H+2][W+2][8]:
 *   element (n,c,y,x) lives at  ((c/8)*plane_stride + n*(H+2)*(W+2) + (y+1)*(W+2) + (x+1))*8 + c%8
 * with a zero 1-pixel halo per crop that kernels never write.  The caller must keep
 * POCO_ACT_GUARD_BYTES of readable (zeroed) memory before `data` and after the last plane.
 */
#ifndef POCO_B200_H
#define POCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POCO_ACT_GUARD_BYTES 8192
#define POCO_MAX_FUSE_INPUTS 4

typedef struct poco_act {
    void* data;           /* fp16, plane 0 / padded pixel 0 */
    int64_t plane_stride; /* pixels per plane: N*(H+2)*(W+2) */
    int32_t C, N, H, W;   /* channels (multiple of 8), crops, unpadded height / width */
    void* lo;             /* NULL: fp16 mode.  Split-precision ("parity") mode: a second fp16 tensor of the same
                           * geometry (same plane_stride, zero halo, guards) holding the rounding residual, so that
                           * the value of an element is  float(data[i]) + float(lo[i])  (~22 significant bits).
                           * Every op that reads / writes a poco_act honours it; all tensors of one op must agree. */
} poco_act;

/* conv + folded BatchNorm + optional residual add + optional ReLU.
 * Replaces nn.Conv2d -> nn.BatchNorm2d(eval) -> (+= residual) -> nn.ReLU chains, e.g.
 * BasicBlock/Bottleneck (backbone/hrnet.py:42-58, :79-99; resnet.py:100-121), transition and fuse
 * convs (hrnet.py:198-240, :345-384), output-stage convs (hrnet.py:437-450), the PARE conv branches
 * (head/pare_head.py:468-491).  weight: fp16 [kh*kw][Cin/8][Cout][8] with the BN scale folded in,
 * bias: fp32 [Cout] (BN shift + conv bias).  Cin, Cout multiples of 16.
 * Split-precision mode (in.lo != NULL; the reference computes in fp32, pocolib/core/config.py:154 PRECISION=32):
 * weight holds TWO such tensors back to back, W_hi = fp16(W) then W_lo = fp16(W - W_hi), and the kernel accumulates
 * x_hi W_hi + x_lo W_hi + x_hi W_lo in one fp32 MMA chain (3x the tensor work, operands good to ~2^-22), adds
 * residual + residual_lo in fp32 and stores out = fp16(y), out.lo = fp16(y - out). */
typedef struct poco_conv {
    poco_act in, out;
    const void* weight;
    const float* bias;
    const void* residual; /* fp16, same geometry as out (or NULL) */
    int64_t res_plane_stride;
    int32_t kh, kw, stride, pad;
    int32_t relu; /* 0 none; 1 ReLU after the residual add; 2 ReLU before the residual add */
    int32_t impl; /* 0 = tcgen05 implicit GEMM (product path); 1 = CUDA-core debug kernel */
    int32_t max_ctas; /* 0 = one CTA per SM; else cap on the persistent grid (concurrent plan lanes share the SMs) */
    int32_t wfmt; /* weight layout: 0 = [kh*kw][Cin/8][Cout][8]; 1 = "dx in N" for 3x3 / stride 1 / pad 1 with
                   * Cout in {32, 64}: [kh][Cin/8][kw*Cout][8] (column s*Cout + co of filter row r) -- the three
                   * horizontal taps share one MMA, a third of the shared-memory operand reads;
                   * 2 = split precision with N-concatenated weights, stride-1 3x3 / 1x1 convs with Cout <= 64:
                   * ONE tensor [kh*kw][Cin/8][2*Cout][8] whose slab rows are W_hi (rows 0..Cout-1) then W_lo -- x_hi meets
                   * both in one N = 2*Cout MMA, the two column groups are summed in the epilogue */
    const void* residual_lo; /* split-precision mode: rounding residual of `residual` (same plane stride), else NULL */
    /* Space-to-depth plumbing of the stride-2 convs (hrnet.py:213-240, :345-384, resnet.py:100-121 with stride 2).
     * A 3x3 / stride 2 / pad 1 conv reads input pixel (2y + r - 1, 2x + s - 1): in the PHASE-SPLIT form of its input --
     * 4*C channels at half resolution, channel ((y & 1) * 2 + (x & 1)) * C + c of pixel (y / 2, x / 2) -- every tap is
     * one phase block at a shift of -1 / 0 rows and pixels, so the conv runs on the contiguous halo-run path of the
     * stride-1 convs (one bulk copy per plane, every input byte fetched once) instead of gathering 16 bytes of every
     * 32-byte sector.  The producing conv writes that form from its epilogue: */
    poco_act out_s2d;  /* data NULL: none.  Else a second output, the phase-split form of `out` (4*out.C channels,
                        * out.H/2 x out.W/2, H and W even, same precision mode), written next to `out` */
    int32_t s2d_only;  /* 1: write out_s2d only (`out` still describes the geometry; its memory is not touched) */
    int32_t in_s2d;    /* 1: `in` IS the phase-split form of the real input (in.C = 4*Cin, in.H = out.H) and the conv is
                        * the 3x3 / stride 2 / pad 1 conv of that real input; weights keep the [9][Cin/8][Cout][8] layout */
} poco_conv;

/* A chain of convolutions of ONE geometry (3x3/s1/p1 or 1x1/s1, Cin == Cout) executed by one persistent
 * launch: segment i reads what segment i-1 wrote (seg[i].in.data == seg[i-1].out.data).  Used for the
 * four BasicBlocks of an HRNet branch (hrnet.py:42-58 x4 inside HighResolutionModule._make_one_branch,
 * hrnet.py:140-186): eight convs, one launch.  Tiles of segment i+1 start as soon as the three tiles of
 * segment i they read are complete (per-tile counters in `flags`), so there is no grid-wide barrier and
 * no per-conv launch / pipeline fill.  A residual must come from outside the chain or from a segment
 * at least two before its use.  flags: (n_seg - 1) * ceil(N*(H+2)*(W+2) / 128) int32, owned by the
 * caller, zeroed by the call itself (cudaMemsetAsync on `stream`) before the kernel. */
#define POCO_MAX_CHAIN 8
typedef struct poco_conv_chain {
    poco_conv seg[POCO_MAX_CHAIN];
    int32_t n_seg;
    int32_t pad_;
    int32_t* flags;
} poco_conv_chain;

/* One residual BasicBlock as ONE launch (hrnet.py:42-58 without a downsample branch):
 *     out = ReLU(BN2(conv2(ReLU(BN1(conv1(in))))) + in),   both convs 3x3 / stride 1 / pad 1, C -> C channels.
 * The intermediate tensor never leaves the SM: a work unit computes the conv1 tiles that cover its conv2
 * tiles plus their (W+3)-pixel reach (halo recompute), the first epilogue writes them as fp16 into shared
 * memory in the operand layout, and conv2's MMAs read them there; the block input is the residual.
 * C = 32 or 64 with W + 3 <= 64 (poco_basic_block_supported), fp16 mode only (in.lo == out.lo == NULL); `in` and
 * `out` must not alias.
 * weight1 / weight2: poco_conv weight format 0 with the BN scale folded, bias1 / bias2: the BN shifts. */
typedef struct poco_basic_block {
    poco_act in;
    poco_act out;
    const void* weight1;
    const float* bias1;
    const void* weight2;
    const float* bias2;
    int32_t max_ctas; /* like poco_conv.max_ctas: 0 = all SMs */
    int32_t pad_;
    poco_act out_s2d; /* data NULL: none.  Else a second output like poco_conv.out_s2d: the phase-split form of `out`
                       * (4*C channels, H/2 x W/2, H and W even) for the stride-2 fuse convs that read this block
                       * (hrnet.py:213-240); C = 32 only */
} poco_basic_block;

/* The tail of a Bottleneck block as ONE launch (hrnet.py:79-99, resnet.py:100-121):
 *     out = ReLU(BN3(conv3(ReLU(BN2(conv2(in))))) + residual),   conv2 3x3 / stride 1 / pad 1 (Cmid -> Cmid),
 * conv3 1x1 (Cmid -> Cout).  conv2's output tile goes through shared memory straight into conv3's MMAs (a 1x1 conv
 * is pointwise: no halo, no recompute).  Cmid = 64, Cout = 256 (layer1 of the HRNet trunks), fp16 mode only;
 * `residual` has `out`'s geometry (plane stride res_plane_stride pixels) and must not alias `out`.
 * weight2 / weight3: poco_conv weight format 0 with the BN scales folded, bias2 / bias3: the BN shifts. */
typedef struct poco_bottleneck_tail {
    poco_act in;
    poco_act out;
    const void* residual;
    int64_t res_plane_stride;
    const void* weight2;
    const float* bias2;
    const void* weight3;
    const float* bias3;
    int32_t max_ctas; /* like poco_conv.max_ctas: 0 = all SMs */
    int32_t pad_;
} poco_bottleneck_tail;

/* The BasicBlocks of one low-resolution HRNet branch as ONE launch with the crop resident in shared memory
 * (hrnet.py:42-58 x n_blocks inside HighResolutionModule._make_one_branch, hrnet.py:140-186):
 *     x <- ReLU(BN2(conv2(ReLU(BN1(conv1(x))))) + x)   n_blocks times,   every conv 3x3 / stride 1 / pad 1, C -> C.
 * A work unit is one crop: its padded image is at most two 128-pixel tiles and all C channels fit in shared memory, so
 * the activations of all 2 * n_blocks convs never leave the SM (no halo recompute, no HBM hand-off between convs); only
 * the weights stream out of L2.  C = 128 and (H + 2) * (W + 2) <= 256 (poco_branch_supported: the 14x14 branch of
 * HRNet-W32), fp16 mode only (in.lo == out.lo == NULL); `out` may alias `in`.
 * weight[2b] / weight[2b + 1]: conv1 / conv2 of block b in poco_conv weight format 0 with the BN scale folded,
 * bias[...]: the BN shifts. */
#define POCO_MAX_BRANCH_BLOCKS 4
typedef struct poco_branch {
    poco_act in;
    poco_act out;
    const void* weight[2 * POCO_MAX_BRANCH_BLOCKS];
    const float* bias[2 * POCO_MAX_BRANCH_BLOCKS];
    int32_t n_blocks;
    int32_t max_ctas; /* like poco_conv.max_ctas: 0 = all SMs */
} poco_branch;

/* batch['img'] f32 NCHW [N,3,H,W] -> planar-8 fp16 with channels padded to 16 (poco.py:100 input) */
typedef struct poco_pack_image {
    const float* img;
    poco_act out;
    /* 0: out = [16 ch][N][H][W], channels 3..15 zero.
     * 1: im2col of a 3x3 / stride 2 / pad 1 window (the HRNet stem conv1, hrnet.py:299-301, :467-469):
     *    img is [N,3,2*out.H,2*out.W], out = [32 ch][N][out.H][out.W] with channel (r*3+s)*3+c =
     *    img[n, c, 2y+r-1, 2x+s-1] (0 outside), channels 27..31 zero -- the stem conv then runs as a 1x1 conv
     *    with K = 32 instead of a 3x3 conv over 16 zero-padded channels (half the activation bytes). */
    int32_t im2col;
    int32_t pad_;
} poco_pack_image;

/* out = relu?( sum_k nearest_upsample(in[k], 2^shift[k]) ): the HRNet multi-resolution fuse
 * (hrnet.py:257-264 with the nn.Upsample(mode='nearest') of :206-208 folded into the index) */
typedef struct poco_fuse_sum {
    poco_act out;
    poco_act in[POCO_MAX_FUSE_INPUTS];
    int32_t shift[POCO_MAX_FUSE_INPUTS];
    int32_t n_in;
    int32_t relu;
} poco_fuse_sum;

/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (hrnet.py:440) */
typedef struct poco_upsample2x {
    poco_act in, out;
} poco_upsample2x;

/* nn.MaxPool2d(3, 2, 1) (resnet.py:156, :205) */
typedef struct poco_maxpool {
    poco_act in, out;
} poco_maxpool;

/* global average pool -> f32 [N, C] with leading dimension ld (cliff_head.py:95-97,
 * hrnet_cls.py:479-482) */
typedef struct poco_avgpool {
    poco_act in;
    float* out;
    int64_t ld;
} poco_avgpool;

/* planar-8 fp16 -> f32 NCHW [N, C_valid, H, W] (feature export / tests) */
typedef struct poco_unpack {
    poco_act in;
    float* out;
    int32_t c_valid;
} poco_unpack;

/* y = act(x W^T + b) (+ res), fp32 row-major.  Replaces nn.Linear call sites: cliff_head.py:103-113,
 * poco_head.py:122-141, nf_head.py:82, pare_head.py:905-906.  act: 0 none, 1 sigmoid. */
typedef struct poco_linear {
    const float* x;
    int64_t ldx;
    const float* w; /* [O][I] */
    const float* b; /* [O] or NULL */
    const float* res;
    int64_t ldres;
    float* y;
    int64_t ldy;
    int32_t M, I, O;
    int32_t act;
    float* scratch;          /* optional split-K workspace (or NULL): partial sums [splits][M][O] */
    int64_t scratch_floats;  /* its capacity; the library picks splits <= scratch_floats / (M*O) */
} poco_linear;

/* dst[r, c] = src[(bcast ? 0 : r), c]  (torch.cat / .expand plumbing of cliff_head.py:85-101) */
typedef struct poco_copy2d {
    const float* src;
    int64_t lds;
    float* dst;
    int64_t ldd;
    int32_t rows, cols;
    int32_t bcast;
} poco_copy2d;

/* geometry.rot6d_to_rotmat (utils/geometry.py:247-261): x [rows*24? , 6] with row stride ldx ->
 * out [n, 3, 3] contiguous.  n = number of 6-vectors; vector i starts at x + (i / per_row) * ldx +
 * (i % per_row) * 6. */
typedef struct poco_rot6d {
    const float* x;
    int64_t ldx;
    int32_t per_row;
    int32_t n;
    float* out;
} poco_rot6d;

/* PARE head after the two conv branches (pare_head.py:754-826, :896-906, keypoint_attention.py:34-56,
 * locallyconnected2d.py:27-37):
 *   segm = conv1x1(part_feats) (+bias)           -> pred_segm_mask f32 [N,25,H,W]
 *   attn = softmax_HW(segm[:,1:])                 (24 joints)
 *   point_local[c,j] = sum_p attn[j,p] smpl_feats[c,p]      -> uncert_feat [N,128*24]
 *   cam_shape[:,j]   = W_sf point_local[:,j] + b_sf  (the 1x1 smpl_final_layer commutes with the pooling)
 *   pose6d[j,:] = LC(point_local), shape/cam = Linear(flatten(cam_shape)); rot6d -> pred_pose */
typedef struct poco_pare_head {
    poco_act part_feats, smpl_feats; /* 128 channels each */
    const float* w_kp;   /* [25][128] */
    const float* b_kp;   /* [25] */
    const float* w_sf;   /* [64][128] */
    const float* b_sf;   /* [64] */
    const float* w_pose; /* [6][128][24] */
    const float* w_shape; /* [10][1536] */
    const float* b_shape;
    const float* w_cam; /* [3][1536] */
    const float* b_cam;
    float* segm;        /* [N,25,H,W] */
    float* uncert_feat; /* [N,3072] */
    float* pose6d;      /* [N,24,6] */
    float* rotmat;      /* [N,24,3,3] */
    float* shape;       /* [N,10] */
    float* cam;         /* [N,3] */
    float* scratch;     /* poco_pare_scratch_floats(N,H,W) floats */
} poco_pare_head;

/* conditional RealNVP (layers/real_nvp.py:25-65; nets nf_head.py:13-17), fp32.
 * params: packed by poco_b200.engine.pack_realnvp -- per coupling layer i: mask[D],
 * then for net in (s, t): W0[H][D+CTX] b0[H] W1[H][H] b1[H] W2[D][H] b2[D].
 * direction 0: log_prob (inverse pass, out = logp[R], optional z_out[R,D], logdet_out[R]);
 * direction 1: forward_p (sampling pass, out = x[R,D]). */
typedef struct poco_realnvp {
    const float* x;   /* [R, D] */
    const float* ctx; /* [R, CTX] (or NULL when CTX == 0) */
    const float* params;
    float* out;
    float* z_out;
    float* logdet_out;
    int32_t R, D, CTX, HID, L;
    int32_t direction;
    /* Context hoist (optional; ctx_part == NULL: every CTA walks the full [H][D+CTX] first layers itself).  The context
     * columns of the first layer of every (coupling layer, s/t net) do not depend on z, and the 24 joints of a crop
     * share one context vector (nf_head.py:85-101 expands it), so the host computes
     *     ctx_part[g][(layer*2 + net)*H + u] = b0[u] + sum_k W0[u][D + k] * ctx[g][k]
     * ONCE per context row with one GEMM (poco_linear on the matrix pack_realnvp_ctx stacks) and passes it here:
     * row r uses context row r / ctx_group, the kernel only adds the D-column part. */
    const float* ctx_part; /* [ceil(R / ctx_group)][L * 2 * HID] or NULL */
    int32_t ctx_group;     /* rows per context row (1 = one per row) */
    int32_t pad_;
} poco_realnvp;

/* Per-detection crop + normalisation, the step right before the hot path.  Replaces
 * get_single_image_crop_demo (utils/vibe_image_utils.py:233-267: gen_trans_from_patch_cv :58-91, cv2.warpAffine
 * INTER_LINEAR / BORDER_CONSTANT :104-105, ToTensor + Normalize :343-352) and calculate_bbox_info /
 * calculate_focal_length (utils/image_utils.py:171-187) as used by POCOTester (core/tester.py:181-212).
 * frame: uint8 RGB [H][W][3]; boxes: f32 [n][4] = (cx, cy, w, h) in pixels; scale: the bbox scale of the caller
 * (1.2 in the demo).  img: f32 [n][3][crop][crop], bit-identical to the reference (cv2's fixed-point bilinear
 * is reproduced exactly).  The five per-detection outputs may be NULL. */
typedef struct poco_crop {
    const uint8_t* frame;
    int32_t frame_h, frame_w;
    const float* boxes;
    int32_t n, crop;
    float scale;
    int32_t pad_;
    float* img;
    float* bbox_info;    /* [n][3] */
    float* focal_length; /* [n] */
    float* scale_out;    /* [n] max(w, h) / 200 */
    float* center;       /* [n][2] */
    float* orig_shape;   /* [n][2] (h, w) */
} poco_crop;

/* Uncertainty post-processing, the step right after the hot path: POCOUtils.prepare_uncert
 * (utils/poco_utils.py:63-94, with get_kinematic_uncert :21-25 and the `1 - var` confidence option) followed by
 * POCOUtils.get_global_uncert (:50-61), as core/tester.py:243-245 / :418-421 call them.  var: f32 [n][24]
 * (`var_pose` of POCO.forward).  Outputs (each may be NULL): prepared [n][24], thresholded [n][24] (the copy
 * get_global_uncert modifies in place), global_var [n].  cliff != 0: threshold 2 * sensitivity_threshold and
 * global = first entry; else threshold sensitivity_threshold and global = mean of the row. */
typedef struct poco_uncert_post {
    const float* var;
    int32_t n;
    int32_t cliff;
    int32_t kinematic;
    int32_t return_conf;
    float sensitivity_threshold; /* 0.40 in the reference */
    int32_t pad_;
    float* prepared;
    float* thresholded;
    float* global_var;
} poco_uncert_post;

/* SMPL mesh stage, the last step of POCO.forward: smplx.SMPL linear blend skinning with
 * pose2rot=False (smplx==0.1.28 lbs.py, called at models/head/smpl_head.py:53-58 / smplcam_head.py:48-53), the
 * wrapper's vertex joints + J_regressor_extra joints + joint_map (smpl_head.py:12-34), the camera conversions
 * (utils/geometry.py:447-463, smplcam_head.py:123-139) and the 2-D projection (geometry.py:480-508,
 * smplcam_head.py:99-120).  smplx and the licensed model files are absent from the reference tree: the LBS part
 * restates the published algorithm and its parity is unpinned.
 * The model is device data prepared once by the host (poco_b200/smpl.py prepare_smpl_model): vertex arrays are
 * coordinate-major and padded to vp = nv rounded up to 128 (padding holds zeros). */
#define POCO_SMPL_JOINTS 24
#define POCO_SMPL_BETAS 10
#define POCO_SMPL_DIR_ROWS 224 /* blend table rows: 10 shape + 207 pose directions + 7 zero rows */
#define POCO_SMPL_SCRATCH_FLOATS 580 /* per crop: 220 blend coefficients, 24x12 transforms, 24x3 posed joints */
typedef struct poco_smpl_model {
    const float* v_template;         /* [3][vp] */
    const float* dirs;               /* [224][3][vp]: rows 0..9 shapedirs, 10..216 posedirs, 217..223 zero */
    const float* weights;            /* [24][vp] skinning weights */
    const float* j_template;         /* [24][3]      J_regressor . v_template */
    const float* j_dirs;             /* [24][3][10]  J_regressor . shapedirs  */
    const int32_t* parents;          /* [24] kinematic tree, parents[0] = -1, parents[i] < i */
    const int32_t* extra_vertex_ids; /* [n_extra_vertex] smplx VertexJointSelector */
    const int32_t* reg_row_ptr;      /* CSR of J_regressor_extra: [n_extra_reg + 1] */
    const int32_t* reg_col;
    const float* reg_val;
    const int32_t* joint_map;        /* [n_joints_out] into the 24 + n_extra_vertex + n_extra_reg joints (<= 64) */
    int32_t nv, vp, n_extra_vertex, n_extra_reg, n_joints_out, pad_;
} poco_smpl_model;

typedef struct poco_smpl {
    poco_smpl_model model;
    const float* rotmat; /* [n][24][3][3] pred_pose */
    const float* betas;  /* [n][10] pred_shape */
    const float* cam;    /* [n][3] pred_cam (s, tx, ty) */
    const float* focal_length; /* cliff only: [n] */
    const float* bbox_scale;   /* [n] bbox height / 200 */
    const float* bbox_center;  /* [n][2] */
    const float* img_w;        /* [n] */
    const float* img_h;        /* [n] */
    int32_t n;
    int32_t cliff;              /* 0: smpl_head (crop camera, focal_default, centre 0); 1: smplcam_head */
    int32_t normalize_joints2d; /* joints2d / (img_res / 2) (smpl_head.py:77-79) */
    int32_t img_res;
    float focal_default; /* 5000 */
    int32_t pad_;
    float* scratch;       /* n * POCO_SMPL_SCRATCH_FLOATS floats */
    float* vertices;      /* [n][nv][3] */
    float* joints3d;      /* [n][n_joints_out][3] */
    float* joints2d;      /* [n][n_joints_out][2] or NULL */
    float* cam_t;         /* [n][3] or NULL */
    float* fullimg_cam_t; /* [n][3] or NULL (cliff) */
} poco_smpl;

/* fork / join of plan lanes.  HRNet's branches (and the per-output fuse chains) are independent, and the
 * low-resolution ones cannot fill 148 SMs on their own: a plan runs them concurrently on internal
 * streams (lane k of `poco_op.lane`), each conv capped to its share of the SMs (poco_conv.max_ctas). */
typedef struct poco_sync {
    int32_t n_lanes;
} poco_sync;

typedef enum poco_op_kind {
    POCO_OP_PACK_IMAGE = 1,
    POCO_OP_CONV = 2,
    POCO_OP_FUSE_SUM = 3,
    POCO_OP_UPSAMPLE2X = 4,
    POCO_OP_MAXPOOL = 5,
    POCO_OP_AVGPOOL = 6,
    POCO_OP_UNPACK = 7,
    POCO_OP_LINEAR = 8,
    POCO_OP_COPY2D = 9,
    POCO_OP_ROT6D = 10,
    POCO_OP_PARE_HEAD = 11,
    POCO_OP_REALNVP = 12,
    POCO_OP_FORK = 13, /* lanes 1..n-1 start after everything enqueued so far on lane 0 */
    POCO_OP_JOIN = 14, /* lane 0 continues after lanes 1..n-1 have drained */
    POCO_OP_CONV_CHAIN = 15,
    POCO_OP_CROP = 16,
    POCO_OP_UNCERT_POST = 17,
    POCO_OP_SMPL = 18,
    POCO_OP_BASIC_BLOCK = 19,
    POCO_OP_BOTTLENECK_TAIL = 20,
    POCO_OP_BRANCH = 21
} poco_op_kind;

typedef struct poco_op {
    int32_t kind;
    int32_t lane; /* execution lane (stream) inside a plan; 0 = main */
    union {
        poco_pack_image pack_image;
        poco_conv conv;
        poco_conv_chain conv_chain;
        poco_basic_block basic_block;
        poco_bottleneck_tail bottleneck_tail;
        poco_branch branch;
        poco_fuse_sum fuse_sum;
        poco_upsample2x upsample2x;
        poco_maxpool maxpool;
        poco_avgpool avgpool;
        poco_unpack unpack;
        poco_linear linear;
        poco_copy2d copy2d;
        poco_rot6d rot6d;
        poco_pare_head pare_head;
        poco_realnvp realnvp;
        poco_sync sync;
        poco_crop crop;
        poco_uncert_post uncert_post;
        poco_smpl smpl;
    } u;
} poco_op;

typedef struct poco_plan poco_plan;

/* library / device */
int poco_version(void);
const char* poco_last_error(void);
int poco_device_check(int device); /* 0 iff `device` is an sm_100 part */
int64_t poco_kernel_launches(void); /* kernels launched by this library since load (bench evidence) */

/* single ops (each equals poco_run_op on the matching poco_op) */
int poco_run_op(const poco_op* op, void* stream);
int poco_conv_run(const poco_conv* d, void* stream);
int poco_conv_chain_run(const poco_conv_chain* d, void* stream);
int poco_basic_block_run(const poco_basic_block* d, void* stream);
int poco_basic_block_supported(int32_t C, int32_t H, int32_t W); /* 1 iff poco_basic_block_run takes this geometry */
int poco_bottleneck_tail_run(const poco_bottleneck_tail* d, void* stream);
int poco_bottleneck_tail_supported(int32_t Cmid, int32_t Cout, int32_t H, int32_t W);
int poco_branch_run(const poco_branch* d, void* stream);
int poco_branch_supported(int32_t C, int32_t H, int32_t W, int32_t n_blocks); /* 1 iff poco_branch_run takes this geometry */
int64_t poco_conv_chain_flag_count(const poco_conv_chain* d); /* int32 entries `flags` must hold */
int poco_pack_image_run(const poco_pack_image* d, void* stream);
int poco_fuse_sum_run(const poco_fuse_sum* d, void* stream);
int poco_upsample2x_run(const poco_upsample2x* d, void* stream);
int poco_maxpool_run(const poco_maxpool* d, void* stream);
int poco_avgpool_run(const poco_avgpool* d, void* stream);
int poco_unpack_run(const poco_unpack* d, void* stream);
int poco_linear_run(const poco_linear* d, void* stream);
int poco_copy2d_run(const poco_copy2d* d, void* stream);
int poco_rot6d_run(const poco_rot6d* d, void* stream);
int poco_pare_head_run(const poco_pare_head* d, void* stream);
int poco_realnvp_run(const poco_realnvp* d, void* stream);
int poco_crop_run(const poco_crop* d, void* stream);
int poco_uncert_post_run(const poco_uncert_post* d, void* stream);
int poco_smpl_run(const poco_smpl* d, void* stream);
int64_t poco_pare_scratch_floats(int32_t N, int32_t H, int32_t W);

/* a plan = the static layer schedule of one POCO.forward for one batch size (poco.py:99-129):
 * ops are validated once at creation and replayed in order by poco_plan_run. */
int poco_plan_create(const poco_op* ops, int32_t n_ops, poco_plan** out);
int poco_plan_run(poco_plan* plan, void* stream);
int32_t poco_plan_num_ops(const poco_plan* plan);
int64_t poco_plan_flops(const poco_plan* plan); /* 2*MAC of conv/linear ops (algorithmic) */
void poco_plan_destroy(poco_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* POCO_B200_H */
