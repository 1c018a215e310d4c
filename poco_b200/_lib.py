"""ctypes binding of libpoco_b200.so (the C ABI declared in include/poco_b200.h).

The product path has NO fallback: if the shared library is missing or an entry point fails, an
exception is raised.  Structures below mirror include/poco_b200.h field by field.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('POCO_B200_LIB') or os.path.join(_HERE, 'libpoco_b200.so')     # (override: A/B benchmarking of builds)

MAX_FUSE_INPUTS = 4
ACT_GUARD_BYTES = 8192

OP_PACK_IMAGE, OP_CONV, OP_FUSE_SUM, OP_UPSAMPLE2X, OP_MAXPOOL, OP_AVGPOOL, OP_UNPACK, OP_LINEAR, \
    OP_COPY2D, OP_ROT6D, OP_PARE_HEAD, OP_REALNVP, OP_FORK, OP_JOIN, OP_CONV_CHAIN, OP_CROP, OP_UNCERT_POST, OP_SMPL, \
    OP_BASIC_BLOCK, OP_BOTTLENECK_TAIL, OP_BRANCH = range(1, 22)
MAX_CHAIN = 8
MAX_BRANCH_BLOCKS = 4
SMPL_JOINTS, SMPL_BETAS, SMPL_SCRATCH_FLOATS, SMPL_DIR_ROWS = 24, 10, 580, 224


class Act(C.Structure):
    _fields_ = [('data', C.c_void_p), ('plane_stride', C.c_int64),
                ('C', C.c_int32), ('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('lo', C.c_void_p)]


class Conv(C.Structure):
    _fields_ = [('in_', Act), ('out', Act), ('weight', C.c_void_p), ('bias', C.c_void_p),
                ('residual', C.c_void_p), ('res_plane_stride', C.c_int64),
                ('kh', C.c_int32), ('kw', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32),
                ('relu', C.c_int32), ('impl', C.c_int32), ('max_ctas', C.c_int32), ('wfmt', C.c_int32),
                ('residual_lo', C.c_void_p), ('out_s2d', Act), ('s2d_only', C.c_int32), ('in_s2d', C.c_int32)]


class ConvChain(C.Structure):
    _fields_ = [('seg', Conv * MAX_CHAIN), ('n_seg', C.c_int32), ('pad_', C.c_int32), ('flags', C.c_void_p)]


class BasicBlock(C.Structure):
    _fields_ = [('in_', Act), ('out', Act), ('weight1', C.c_void_p), ('bias1', C.c_void_p), ('weight2', C.c_void_p),
                ('bias2', C.c_void_p), ('max_ctas', C.c_int32), ('pad_', C.c_int32), ('out_s2d', Act)]


class BottleneckTail(C.Structure):
    _fields_ = [('in_', Act), ('out', Act), ('residual', C.c_void_p), ('res_plane_stride', C.c_int64),
                ('weight2', C.c_void_p), ('bias2', C.c_void_p), ('weight3', C.c_void_p), ('bias3', C.c_void_p),
                ('max_ctas', C.c_int32), ('pad_', C.c_int32)]


class Branch(C.Structure):
    _fields_ = [('in_', Act), ('out', Act), ('weight', C.c_void_p * (2 * MAX_BRANCH_BLOCKS)),
                ('bias', C.c_void_p * (2 * MAX_BRANCH_BLOCKS)), ('n_blocks', C.c_int32), ('max_ctas', C.c_int32)]


class PackImage(C.Structure):
    _fields_ = [('img', C.c_void_p), ('out', Act), ('im2col', C.c_int32), ('pad_', C.c_int32)]


class FuseSum(C.Structure):
    _fields_ = [('out', Act), ('in_', Act * MAX_FUSE_INPUTS), ('shift', C.c_int32 * MAX_FUSE_INPUTS),
                ('n_in', C.c_int32), ('relu', C.c_int32)]


class Upsample2x(C.Structure):
    _fields_ = [('in_', Act), ('out', Act)]


class MaxPool(C.Structure):
    _fields_ = [('in_', Act), ('out', Act)]


class AvgPool(C.Structure):
    _fields_ = [('in_', Act), ('out', C.c_void_p), ('ld', C.c_int64)]


class Unpack(C.Structure):
    _fields_ = [('in_', Act), ('out', C.c_void_p), ('c_valid', C.c_int32)]


class Linear(C.Structure):
    _fields_ = [('x', C.c_void_p), ('ldx', C.c_int64), ('w', C.c_void_p), ('b', C.c_void_p),
                ('res', C.c_void_p), ('ldres', C.c_int64), ('y', C.c_void_p), ('ldy', C.c_int64),
                ('M', C.c_int32), ('I', C.c_int32), ('O', C.c_int32), ('act', C.c_int32),
                ('scratch', C.c_void_p), ('scratch_floats', C.c_int64)]


class Copy2d(C.Structure):
    _fields_ = [('src', C.c_void_p), ('lds', C.c_int64), ('dst', C.c_void_p), ('ldd', C.c_int64),
                ('rows', C.c_int32), ('cols', C.c_int32), ('bcast', C.c_int32)]


class Rot6d(C.Structure):
    _fields_ = [('x', C.c_void_p), ('ldx', C.c_int64), ('per_row', C.c_int32), ('n', C.c_int32),
                ('out', C.c_void_p)]


class PareHead(C.Structure):
    _fields_ = [('part_feats', Act), ('smpl_feats', Act),
                ('w_kp', C.c_void_p), ('b_kp', C.c_void_p), ('w_sf', C.c_void_p), ('b_sf', C.c_void_p),
                ('w_pose', C.c_void_p), ('w_shape', C.c_void_p), ('b_shape', C.c_void_p),
                ('w_cam', C.c_void_p), ('b_cam', C.c_void_p),
                ('segm', C.c_void_p), ('uncert_feat', C.c_void_p), ('pose6d', C.c_void_p),
                ('rotmat', C.c_void_p), ('shape', C.c_void_p), ('cam', C.c_void_p), ('scratch', C.c_void_p)]


class RealNVP(C.Structure):
    _fields_ = [('x', C.c_void_p), ('ctx', C.c_void_p), ('params', C.c_void_p), ('out', C.c_void_p),
                ('z_out', C.c_void_p), ('logdet_out', C.c_void_p),
                ('R', C.c_int32), ('D', C.c_int32), ('CTX', C.c_int32), ('HID', C.c_int32), ('L', C.c_int32),
                ('direction', C.c_int32), ('ctx_part', C.c_void_p), ('ctx_group', C.c_int32), ('pad_', C.c_int32)]


class Crop(C.Structure):
    _fields_ = [('frame', C.c_void_p), ('frame_h', C.c_int32), ('frame_w', C.c_int32), ('boxes', C.c_void_p),
                ('n', C.c_int32), ('crop', C.c_int32), ('scale', C.c_float), ('pad_', C.c_int32), ('img', C.c_void_p),
                ('bbox_info', C.c_void_p), ('focal_length', C.c_void_p), ('scale_out', C.c_void_p),
                ('center', C.c_void_p), ('orig_shape', C.c_void_p)]


class UncertPost(C.Structure):
    _fields_ = [('var', C.c_void_p), ('n', C.c_int32), ('cliff', C.c_int32), ('kinematic', C.c_int32),
                ('return_conf', C.c_int32), ('sensitivity_threshold', C.c_float), ('pad_', C.c_int32),
                ('prepared', C.c_void_p), ('thresholded', C.c_void_p), ('global_var', C.c_void_p)]


class SmplModel(C.Structure):
    _fields_ = [('v_template', C.c_void_p), ('dirs', C.c_void_p), ('weights', C.c_void_p), ('j_template', C.c_void_p),
                ('j_dirs', C.c_void_p), ('parents', C.c_void_p), ('extra_vertex_ids', C.c_void_p),
                ('reg_row_ptr', C.c_void_p), ('reg_col', C.c_void_p), ('reg_val', C.c_void_p), ('joint_map', C.c_void_p),
                ('nv', C.c_int32), ('vp', C.c_int32), ('n_extra_vertex', C.c_int32), ('n_extra_reg', C.c_int32),
                ('n_joints_out', C.c_int32), ('pad_', C.c_int32)]


class Smpl(C.Structure):
    _fields_ = [('model', SmplModel), ('rotmat', C.c_void_p), ('betas', C.c_void_p), ('cam', C.c_void_p),
                ('focal_length', C.c_void_p), ('bbox_scale', C.c_void_p), ('bbox_center', C.c_void_p),
                ('img_w', C.c_void_p), ('img_h', C.c_void_p),
                ('n', C.c_int32), ('cliff', C.c_int32), ('normalize_joints2d', C.c_int32), ('img_res', C.c_int32),
                ('focal_default', C.c_float), ('pad_', C.c_int32),
                ('scratch', C.c_void_p), ('vertices', C.c_void_p), ('joints3d', C.c_void_p), ('joints2d', C.c_void_p),
                ('cam_t', C.c_void_p), ('fullimg_cam_t', C.c_void_p)]


class Sync(C.Structure):
    _fields_ = [('n_lanes', C.c_int32)]


class _OpU(C.Union):
    _fields_ = [('pack_image', PackImage), ('conv', Conv), ('conv_chain', ConvChain), ('basic_block', BasicBlock), ('bottleneck_tail', BottleneckTail),
                ('branch', Branch), ('fuse_sum', FuseSum), ('upsample2x', Upsample2x),
                ('maxpool', MaxPool), ('avgpool', AvgPool), ('unpack', Unpack), ('linear', Linear),
                ('copy2d', Copy2d), ('rot6d', Rot6d), ('pare_head', PareHead), ('realnvp', RealNVP), ('sync', Sync),
                ('crop', Crop), ('uncert_post', UncertPost), ('smpl', Smpl)]


class Op(C.Structure):
    _fields_ = [('kind', C.c_int32), ('lane', C.c_int32), ('u', _OpU)]


_FIELD_OF_KIND = {OP_PACK_IMAGE: 'pack_image', OP_CONV: 'conv', OP_FUSE_SUM: 'fuse_sum',
                  OP_UPSAMPLE2X: 'upsample2x', OP_MAXPOOL: 'maxpool', OP_AVGPOOL: 'avgpool',
                  OP_UNPACK: 'unpack', OP_LINEAR: 'linear', OP_COPY2D: 'copy2d', OP_ROT6D: 'rot6d',
                  OP_PARE_HEAD: 'pare_head', OP_REALNVP: 'realnvp', OP_FORK: 'sync', OP_JOIN: 'sync',
                  OP_CONV_CHAIN: 'conv_chain', OP_CROP: 'crop', OP_UNCERT_POST: 'uncert_post', OP_SMPL: 'smpl',
                  OP_BASIC_BLOCK: 'basic_block', OP_BOTTLENECK_TAIL: 'bottleneck_tail', OP_BRANCH: 'branch'}
_KIND_OF_TYPE = {PackImage: OP_PACK_IMAGE, Conv: OP_CONV, FuseSum: OP_FUSE_SUM, Upsample2x: OP_UPSAMPLE2X,
                 MaxPool: OP_MAXPOOL, AvgPool: OP_AVGPOOL, Unpack: OP_UNPACK, Linear: OP_LINEAR,
                 Copy2d: OP_COPY2D, Rot6d: OP_ROT6D, PareHead: OP_PARE_HEAD, RealNVP: OP_REALNVP,
                 ConvChain: OP_CONV_CHAIN, Crop: OP_CROP, UncertPost: OP_UNCERT_POST, Smpl: OP_SMPL,
                 BasicBlock: OP_BASIC_BLOCK, BottleneckTail: OP_BOTTLENECK_TAIL, Branch: OP_BRANCH}

# every symbol include/poco_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    'poco_version', 'poco_last_error', 'poco_device_check', 'poco_kernel_launches', 'poco_run_op',
    'poco_conv_run', 'poco_conv_chain_run', 'poco_conv_chain_flag_count', 'poco_basic_block_run', 'poco_basic_block_supported',
    'poco_bottleneck_tail_run', 'poco_bottleneck_tail_supported', 'poco_branch_run', 'poco_branch_supported', 'poco_pack_image_run', 'poco_fuse_sum_run', 'poco_upsample2x_run', 'poco_maxpool_run',
    'poco_avgpool_run', 'poco_unpack_run', 'poco_linear_run', 'poco_copy2d_run', 'poco_rot6d_run',
    'poco_pare_head_run', 'poco_realnvp_run', 'poco_crop_run', 'poco_uncert_post_run', 'poco_smpl_run', 'poco_pare_scratch_floats',
    'poco_plan_create', 'poco_plan_run', 'poco_plan_num_ops', 'poco_plan_flops', 'poco_plan_destroy',
]

_lib = None


class PocoError(RuntimeError):
    pass


def lib():
    """Loads libpoco_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PocoError(f'{LIB_PATH} is missing -- build it with `python -c "import __graft_entry__ as g; '
                            f'g.build()"`; poco_b200 has no CPU / eager fallback')
        L = C.CDLL(LIB_PATH)
        L.poco_last_error.restype = C.c_char_p
        L.poco_kernel_launches.restype = C.c_int64
        L.poco_plan_flops.restype = C.c_int64
        L.poco_plan_flops.argtypes = [C.c_void_p]
        L.poco_plan_num_ops.argtypes = [C.c_void_p]
        L.poco_pare_scratch_floats.restype = C.c_int64
        L.poco_pare_scratch_floats.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        if hasattr(L, 'poco_conv_chain_flag_count'):        # (absent only in old builds loaded through POCO_B200_LIB)
            L.poco_conv_chain_flag_count.restype = C.c_int64
            L.poco_conv_chain_flag_count.argtypes = [C.POINTER(ConvChain)]
        if hasattr(L, 'poco_basic_block_supported'):
            L.poco_basic_block_supported.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        if hasattr(L, 'poco_bottleneck_tail_supported'):
            L.poco_bottleneck_tail_supported.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        if hasattr(L, 'poco_branch_supported'):
            L.poco_branch_supported.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.poco_run_op.argtypes = [C.POINTER(Op), C.c_void_p]
        L.poco_plan_create.argtypes = [C.POINTER(Op), C.c_int32, C.POINTER(C.c_void_p)]
        L.poco_plan_run.argtypes = [C.c_void_p, C.c_void_p]
        L.poco_plan_destroy.argtypes = [C.c_void_p]
        L.poco_device_check.argtypes = [C.c_int]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PocoError(lib().poco_last_error().decode())


def make_op(desc, lane=0, kind=None):
    op = Op()
    op.kind = kind if kind is not None else _KIND_OF_TYPE[type(desc)]
    op.lane = lane
    setattr(op.u, _FIELD_OF_KIND[op.kind], desc)
    return op


def run_op(desc, stream):
    """Run one op descriptor on a cudaStream_t handle (int)."""
    op = desc if isinstance(desc, Op) else make_op(desc)
    check(lib().poco_run_op(C.byref(op), C.c_void_p(stream)))


def kernel_launches():
    return int(lib().poco_kernel_launches())
