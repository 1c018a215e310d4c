"""SMPL mesh stage of POCO.forward (reference smpl_head.py:36-83, smplcam_head.py:27-96).

The stage's LBS arithmetic lives in the un-vendored `smplx==0.1.28` package plus licence-gated model files
(data/smpl, data/J_regressor_extra.npy); neither ships with the reference tree.  The stage is therefore *injectable*
and `make_smpl_stage` picks, in this order:

  * `DeviceSmplStage` -- the whole stage (LBS with pose2rot=False, the wrapper's 49 joints, both camera conversions,
    projection) as three CUDA kernels behind one C-ABI call (poco_smpl_run, poco_b200/csrc/smpl.cu), from model arrays
    handed in as data (`POCO(..., smpl_model=...)`; `load_smpl_model` reads an .npz or the official .pkl without
    chumpy).  Needs no smplx and has no CPU path;
  * `SmplStage` -- smplx's own LBS (same calls as the reference) + the torch camera conversions below, when smplx
    and its files are installed but no arrays were given;
  * `StubSmplStage` -- zero meshes with the reference's keys / shapes so POCO.forward keeps its dict contract.

Parity: the camera conversions and the projection are in-tree reference math (pinned by tests/test_smpl.py to
outputs of the reference functions); the LBS restates smplx's published algorithm -- unpinned, excluded from the
1e-3 gate.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

SMPL_MODEL_DIR = 'data/smpl'                              # config.py:38
JOINT_REGRESSOR_TRAIN_EXTRA = 'data/J_regressor_extra.npy'  # config.py:34


def weak_perspective_to_perspective(cam, focal_length=5000., img_res=224):
    """[s, tx, ty] -> [tx, ty, tz]   (geometry.py:447-463)"""
    return torch.stack([cam[:, 1], cam[:, 2], 2 * focal_length / (img_res * cam[:, 0] + 1e-9)], dim=-1)


def crop_cam_to_full_img_cam(cam, bbox_height, bbox_center, img_w, img_h, focal_length, crop_res=224):
    """smplcam_head.convert_pare_to_full_img_cam (smplcam_head.py:123-139)"""
    s, tx, ty = cam[:, 0], cam[:, 1], cam[:, 2]
    r = bbox_height / crop_res
    tz = 2 * focal_length / (r * crop_res * s)
    cx = 2 * (bbox_center[:, 0] - (img_w / 2.)) / (s * bbox_height)
    cy = 2 * (bbox_center[:, 1] - (img_h / 2.)) / (s * bbox_height)
    return torch.stack([tx + cx, ty + cy, tz], dim=-1)


def project(points, translation, K):
    """perspective projection with identity rotation (geometry.py:480-508 / smplcam_head.py:99-120)"""
    p = points + translation.unsqueeze(1)
    p = p / p[:, :, -1:]
    p = torch.einsum('bij,bkj->bki', K, p)
    return p[:, :, :-1]


def _intrinsics(B, device, fx, cx, cy):
    K = torch.zeros(B, 3, 3, device=device)
    K[:, 0, 0] = fx
    K[:, 1, 1] = fx
    K[:, 2, 2] = 1.
    K[:, 0, 2] = cx
    K[:, 1, 2] = cy
    return K


class StubSmplStage(nn.Module):
    """Zero meshes; camera-only outputs are still computed (they do not need SMPL)."""

    def __init__(self, head_name, img_res=224, focal_length=5000.):
        super().__init__()
        self.cliff = 'cliff' in head_name
        self.img_res = img_res
        self.focal_length = focal_length

    def forward(self, rotmat, shape, cam, **kw):
        B, dev = rotmat.shape[0], rotmat.device
        out = {'smpl_vertices': torch.zeros(B, 6890, 3, device=dev),
               'smpl_joints3d': torch.zeros(B, 49, 3, device=dev),
               'smpl_joints2d': torch.zeros(B, 49, 2, device=dev)}
        if self.cliff:
            out['pred_cam_t'] = weak_perspective_to_perspective(cam)
            out['pred_fullimg_cam_t'] = crop_cam_to_full_img_cam(
                cam.detach().clone(), kw['bbox_scale'] * 200., kw['bbox_center'], kw['img_w'], kw['img_h'],
                kw['focal_length'], self.img_res)
        else:
            out['pred_cam_t'] = weak_perspective_to_perspective(cam)
        return out


class SmplStage(nn.Module):
    """Real SMPL LBS through smplx (only constructed when smplx + model files exist)."""

    def __init__(self, head_name, img_res=224, focal_length=5000., joint_map=None):
        super().__init__()
        import numpy as np
        import smplx
        self.cliff = 'cliff' in head_name
        self.img_res = img_res
        self.focal_length = focal_length
        self.smpl = smplx.SMPL(SMPL_MODEL_DIR, create_transl=False)
        self.register_buffer('J_regressor_extra',
                             torch.tensor(np.load(JOINT_REGRESSOR_TRAIN_EXTRA), dtype=torch.float32))
        # [constants.JOINT_MAP[n] for n in constants.JOINT_NAMES] (smpl_head.py:17-20): 49 of the 54 joints
        self.joint_map = list(joint_map) if joint_map is not None else SMPL_JOINT_MAP

    def _joints(self, out):
        extra = torch.einsum('bik,ji->bjk', out.vertices, self.J_regressor_extra)
        j = torch.cat([out.joints, extra], dim=1)
        return j[:, self.joint_map] if self.joint_map is not None else j

    def forward(self, rotmat, shape, cam, normalize_joints2d=False, **kw):
        so = self.smpl(betas=shape.contiguous(), body_pose=rotmat[:, 1:].contiguous(),
                       global_orient=rotmat[:, 0].unsqueeze(1).contiguous(), pose2rot=False)
        joints = self._joints(so)
        B, dev = joints.shape[0], joints.device
        out = {'smpl_vertices': so.vertices, 'smpl_joints3d': joints}
        crop_t = weak_perspective_to_perspective(cam)
        if self.cliff:
            K = _intrinsics(B, dev, kw['focal_length'], kw['img_w'] / 2., kw['img_h'] / 2.)
            full_t = crop_cam_to_full_img_cam(cam.detach().clone(), kw['bbox_scale'] * 200., kw['bbox_center'],
                                              kw['img_w'], kw['img_h'], K[:, 0, 0], self.img_res)
            out['smpl_joints2d'] = project(joints, full_t, K)
            out['pred_cam_t'] = crop_t
            out['pred_fullimg_cam_t'] = full_t
        else:
            K = _intrinsics(B, dev, self.focal_length, 0., 0.)
            j2d = project(joints, crop_t, K)
            if normalize_joints2d:
                j2d = j2d / (self.img_res / 2.)
            out['smpl_joints2d'] = j2d
            out['pred_cam_t'] = crop_t
        return out


# SMPL kinematic tree and the joint bookkeeping of smplx 0.1.28 / the reference wrapper, used when a model file does
# not carry them: vertex joints in VertexJointSelector order (face, feet, left and right finger tips; smplx
# vertex_ids.py 'smplh'), JOINT_MAP = [constants.JOINT_MAP[n] for n in constants.JOINT_NAMES] (constants.py:15-93)
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
SMPL_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                         2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]
SMPL_JOINT_MAP = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                  8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]


def prepare_smpl_model(model):
    """model arrays (smplx naming: v_template [V,3], shapedirs [V,3,>=10], posedirs [207,3V] or the .pkl's [V,3,207],
    J_regressor [24,V], weights [V,24], optional parents / extra_vertex_ids / J_regressor_extra [E,V] / joint_map)
    -> dict of numpy arrays in the layout poco_smpl_model wants (include/poco_b200.h): coordinate-major vertex arrays
    padded to a multiple of 128, shape and pose directions stacked into one [224,3,vp] table (217 rows + zero rows), the joint regressor
    folded through the shape blend (vertices2joints is linear), the extra regressor as CSR."""
    f8 = np.float64
    vt = np.asarray(model['v_template'], f8)
    nv = vt.shape[0]
    sd = np.asarray(model['shapedirs'], f8)[:, :, :L.SMPL_BETAS]
    pd = np.asarray(model['posedirs'], f8)
    n_pose = (L.SMPL_JOINTS - 1) * 9
    if pd.ndim == 3:                                    # official .pkl layout [V,3,207] -> smplx's [207, 3V]
        pd = pd.reshape(-1, pd.shape[-1]).T
    if vt.shape != (nv, 3) or sd.shape != (nv, 3, L.SMPL_BETAS) or pd.shape != (n_pose, nv * 3):
        raise ValueError(f'SMPL model arrays have unexpected shapes: {vt.shape} {sd.shape} {pd.shape}')
    Jr = np.asarray(model['J_regressor'].todense() if hasattr(model['J_regressor'], 'todense') else model['J_regressor'], f8)
    W = np.asarray(model['weights'] if 'weights' in model else model['lbs_weights'], f8)
    if Jr.shape != (L.SMPL_JOINTS, nv) or W.shape != (nv, L.SMPL_JOINTS):
        raise ValueError(f'SMPL J_regressor / weights have unexpected shapes: {Jr.shape} {W.shape}')
    parents = np.asarray(model['parents'] if 'parents' in model else SMPL_PARENTS, np.int64).copy()
    parents[0] = -1
    if parents.shape != (L.SMPL_JOINTS,) or any(parents[i] >= i or parents[i] < 0 for i in range(1, L.SMPL_JOINTS)):
        raise ValueError('SMPL kinematic tree must list parents before children')
    vp = (nv + 127) // 128 * 128

    def planar(a):                                      # [..., V] -> [..., vp] zero padded
        out = np.zeros(a.shape[:-1] + (vp,), np.float32)
        out[..., :nv] = a
        return out
    dirs = np.concatenate([sd.transpose(2, 1, 0), pd.reshape(n_pose, nv, 3).transpose(0, 2, 1),
                           np.zeros((L.SMPL_DIR_ROWS - L.SMPL_BETAS - n_pose, 3, nv))], axis=0)
    ev = np.asarray(model['extra_vertex_ids'] if 'extra_vertex_ids' in model else SMPL_EXTRA_VERTEX_IDS, np.int64)
    if ev.size and (ev.min() < 0 or ev.max() >= nv):
        raise ValueError('extra_vertex_ids out of range')
    Je = np.asarray(model['J_regressor_extra'], f8) if model.get('J_regressor_extra') is not None else np.zeros((0, nv))
    if Je.ndim != 2 or Je.shape[1] != nv:
        raise ValueError(f'J_regressor_extra must be [E, {nv}]')
    rows, cols = np.nonzero(Je)
    row_ptr = np.zeros(Je.shape[0] + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=Je.shape[0]), out=row_ptr[1:])
    n_all = L.SMPL_JOINTS + ev.size + Je.shape[0]
    jm = np.asarray(model['joint_map'] if 'joint_map' in model else SMPL_JOINT_MAP, np.int64)
    if n_all > 64 or jm.size == 0 or jm.min() < 0 or jm.max() >= n_all:
        raise ValueError(f'joint_map must index the {n_all} joints (24 + vertex joints + extra regressor, at most 64)')
    return {'v_template': planar(vt.T), 'dirs': planar(dirs), 'weights': planar(W.T),
            'j_template': (Jr @ vt).astype(np.float32), 'j_dirs': np.einsum('jv,vcl->jcl', Jr, sd).astype(np.float32),
            'parents': parents.astype(np.int32), 'extra_vertex_ids': ev.astype(np.int32),
            'reg_row_ptr': row_ptr, 'reg_col': cols.astype(np.int32), 'reg_val': Je[rows, cols].astype(np.float32),
            'joint_map': jm.astype(np.int32), 'nv': nv, 'vp': vp}


def load_smpl_model(path=None, regressor_extra=JOINT_REGRESSOR_TRAIN_EXTRA):
    """model arrays from an .npz (keys as prepare_smpl_model) or from the official SMPL .pkl under data/smpl
    (config.py:38; read with a chumpy-free unpickler: only the arrays' values are needed)"""
    path = path or os.path.join(SMPL_MODEL_DIR, 'SMPL_NEUTRAL.pkl')
    if path.endswith('.npz'):
        model = dict(np.load(path, allow_pickle=False))
    else:
        import pickle

        class _Array:                                   # stands in for chumpy.Ch objects: keep the ndarray payload
            def __setstate__(self, state):
                self.__dict__.update(state if isinstance(state, dict) else {})

        class _Unpickler(pickle.Unpickler):
            def find_class(self, module, name):
                return _Array if module.startswith('chumpy') else super().find_class(module, name)
        with open(path, 'rb') as fh:
            raw = _Unpickler(fh, encoding='latin1').load()
        model = {k: (v.__dict__.get('x', v.__dict__.get('r')) if isinstance(v, _Array) else v) for k, v in raw.items()}
        if 'kintree_table' in model:
            model['parents'] = np.asarray(model['kintree_table'])[0].astype(np.int64)
    if 'J_regressor_extra' not in model and regressor_extra and os.path.exists(regressor_extra):
        model['J_regressor_extra'] = np.load(regressor_extra)
    return model


class DeviceSmplStage(nn.Module):
    """smpl_head / smplcam_head of the reference (smpl_head.py:36-83, smplcam_head.py:27-96) as one C-ABI op
    (poco_smpl_run: three CUDA kernels).  `model` = arrays as accepted by prepare_smpl_model.  Outputs are fresh
    device tensors with the reference's keys and shapes."""

    def __init__(self, head_name, model, img_res=224, focal_length=5000.):
        super().__init__()
        self.cliff = 'cliff' in head_name
        self.img_res, self.focal_length = int(img_res), float(focal_length)
        p = prepare_smpl_model(model)
        self.nv, self.vp = p.pop('nv'), p.pop('vp')
        self._names = sorted(p)
        for k in self._names:
            self.register_buffer('m_' + k, torch.from_numpy(np.ascontiguousarray(p[k])), persistent=False)
        self.n_joints_out = int(self.m_joint_map.numel())

    def _model_desc(self):
        ptr = {k: getattr(self, 'm_' + k).data_ptr() for k in self._names}
        return L.SmplModel(ptr['v_template'], ptr['dirs'], ptr['weights'], ptr['j_template'], ptr['j_dirs'],
                           ptr['parents'], ptr['extra_vertex_ids'], ptr['reg_row_ptr'], ptr['reg_col'], ptr['reg_val'],
                           ptr['joint_map'], self.nv, self.vp, int(self.m_extra_vertex_ids.numel()),
                           int(self.m_reg_row_ptr.numel()) - 1, self.n_joints_out, 0)

    def forward(self, rotmat, shape, cam, normalize_joints2d=False, stream=None, **kw):
        dev = self.m_dirs.device
        if not (rotmat.is_cuda and dev.type == 'cuda'):
            raise L.PocoError('DeviceSmplStage needs CUDA tensors and a model moved to the GPU (poco_b200 has no CPU path)')
        f32 = lambda t: torch.as_tensor(t, dtype=torch.float32, device=dev).contiguous()     # noqa: E731
        n = rotmat.shape[0]
        rotmat, shape, cam = f32(rotmat).view(n, 24, 3, 3), f32(shape).view(n, 10), f32(cam).view(n, 3)
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)                    # noqa: E731
        out = {'smpl_vertices': new(n, self.nv, 3), 'smpl_joints3d': new(n, self.n_joints_out, 3),
               'smpl_joints2d': new(n, self.n_joints_out, 2), 'pred_cam_t': new(n, 3)}
        scratch = new(n, L.SMPL_SCRATCH_FLOATS)
        cl = [0, 0, 0, 0, 0]
        if self.cliff:
            out['pred_fullimg_cam_t'] = new(n, 3)
            keep = [f32(kw['focal_length']).expand(n) if f32(kw['focal_length']).dim() == 0 else f32(kw['focal_length']).view(n),
                    f32(kw['bbox_scale']).view(n), f32(kw['bbox_center']).view(n, 2), f32(kw['img_w']).view(n),
                    f32(kw['img_h']).view(n)]
            keep = [t.contiguous() for t in keep]
            cl = [t.data_ptr() for t in keep]
        d = L.Smpl(self._model_desc(), rotmat.data_ptr(), shape.data_ptr(), cam.data_ptr(), *cl,
                   n, int(self.cliff), int(bool(normalize_joints2d)), self.img_res, self.focal_length, 0,
                   scratch.data_ptr(), out['smpl_vertices'].data_ptr(), out['smpl_joints3d'].data_ptr(),
                   out['smpl_joints2d'].data_ptr(), out['pred_cam_t'].data_ptr(),
                   out['pred_fullimg_cam_t'].data_ptr() if self.cliff else 0)
        with torch.cuda.device(dev):
            L.run_op(d, stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        return out


def make_smpl_stage(head_name, img_res=224, model=None):
    """model arrays given (or data/smpl/SMPL_NEUTRAL.pkl readable) -> DeviceSmplStage (CUDA); else smplx's LBS when
    that package and its files exist; else the stub (with a warning: its meshes are zeros)."""
    import warnings
    pkl = os.path.join(SMPL_MODEL_DIR, 'SMPL_NEUTRAL.pkl')
    if model is None and os.path.exists(pkl):
        try:
            model = load_smpl_model()
        except Exception as e:                          # noqa: BLE001  (unreadable pickle: say so, then try smplx)
            warnings.warn(f'poco_b200: could not read {pkl} ({type(e).__name__}: {e}); trying smplx', stacklevel=2)
            model = None
    if model is not None:
        return DeviceSmplStage(head_name, model, img_res)
    try:
        import smplx  # noqa: F401
        if os.path.isdir(SMPL_MODEL_DIR) and os.path.exists(JOINT_REGRESSOR_TRAIN_EXTRA):
            return SmplStage(head_name, img_res)
    except Exception:                                   # noqa: BLE001  (smplx absent or its files unreadable)
        pass
    warnings.warn('poco_b200: no SMPL model (pass smpl_model=..., or provide data/smpl + data/J_regressor_extra.npy): '
                  'smpl_vertices / smpl_joints3d / smpl_joints2d are zeros; pose, shape, camera and confidence outputs '
                  'are unaffected', stacklevel=2)
    return StubSmplStage(head_name, img_res)
