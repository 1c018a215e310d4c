"""Host-side SMPL mesh stage of POCO.forward (reference smpl_head.py:36-83, smplcam_head.py:27-96).

Per the north star this stage stays host-side PyTorch.  Its arithmetic lives in the un-vendored
`smplx==0.1.28` package plus licence-gated model files (data/smpl, data/J_regressor_extra.npy), none
of which are available here, so the stage is *injectable*:

  * when `smplx` and the model files are present, `SmplStage` runs the real LBS (same calls as the
    reference) and the camera conversions below;
  * otherwise `StubSmplStage` returns zero meshes with the right keys / shapes so POCO.forward keeps
    its dict contract (parity for smpl_* keys is unpinned and excluded from the 1e-3 gate).

The camera conversions are in-tree reference math and are implemented here in plain torch.
"""
import os

import torch
import torch.nn as nn

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
        self.joint_map = joint_map      # constants.JOINT_MAP order of the reference (49 joints)

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


def make_smpl_stage(head_name, img_res=224):
    try:
        import smplx  # noqa: F401
        if os.path.isdir(SMPL_MODEL_DIR) and os.path.exists(JOINT_REGRESSOR_TRAIN_EXTRA):
            return SmplStage(head_name, img_res)
    except Exception:
        pass
    return StubSmplStage(head_name, img_res)
