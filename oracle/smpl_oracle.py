"""CPU oracle of the SMPL mesh stage of POCO.forward.  TEST INFRASTRUCTURE ONLY: nothing under
poco_b200/ imports this file; the product path is poco_b200/csrc/smpl.cu and fails loudly without its library.

PARITY UNPINNED.  The arithmetic of this stage lives in a third-party dependency that is absent from the reference
tree: `smplx==0.1.28` (reference requirements.txt:7), called at pocolib/models/head/smpl_head.py:12-34, 53-58 and
smplcam_head.py:48-53, with licence-gated model files (config.py:34, 38).  Neither the package nor the model files
are available here and the reference has no test or golden vector at this boundary, so what follows restates the
published algorithm of that version (Loper et al., "SMPL: A Skinned Multi-Person Linear Model", 2015; smplx/lbs.py
`lbs`, `blend_shapes`, `vertices2joints`, `batch_rigid_transform`; smplx/vertex_joint_selector.py) in numpy
float64 and is anchored on the reference's own call sites:

  smplx.SMPL.forward(betas, body_pose, global_orient, pose2rot=False)    smpl_head.py:53-58
  + J_regressor_extra joints, joint_map selection                        smpl_head.py:23-27
  + camera conversions and projection (these ARE in the reference tree)  geometry.py:447-463, 480-508,
                                                                         smplcam_head.py:58-139

The in-tree parts (cameras, projection) are pinned by tests/test_smpl.py against outputs of the reference functions (tests/golden/smpl_cam_golden.npz);
the LBS part is checked through properties of the algorithm (rest pose, rigid motion of the root, blend linearity).
The model is data: any (v_template, shapedirs, posedirs, J_regressor, weights, parents) works, `synthetic_model`
makes a seeded one of the real SMPL dimensions.
"""
import numpy as np

from synth.smpl_model import (EXTRA_VERTEX_IDS, JOINT_MAP, NB, NJ, NV, PARENTS,  # noqa: F401  (data: constants and the
                              synthetic_model)                                     # seeded stand-in model)


def lbs(model, betas, rotmats):
    """smplx.lbs.lbs with pose2rot=False.  betas [B,10], rotmats [B,24,3,3] -> (vertices [B,V,3], joints [B,24,3])"""
    f = np.float64
    vt, sd, pd = model['v_template'].astype(f), model['shapedirs'].astype(f), model['posedirs'].astype(f)
    Jr, W, parents = model['J_regressor'].astype(f), model['weights'].astype(f), model['parents']
    betas, R = np.asarray(betas, f), np.asarray(rotmats, f)
    B = betas.shape[0]
    v_shaped = vt[None] + np.einsum('bl,mkl->bmk', betas, sd)                      # blend_shapes
    J = np.einsum('bik,ji->bjk', v_shaped, Jr)                                     # vertices2joints
    pose_feature = (R[:, 1:] - np.eye(3)).reshape(B, -1)
    v_posed = v_shaped + (pose_feature @ pd).reshape(B, -1, 3)
    # batch_rigid_transform
    rel = J.copy()
    rel[:, 1:] -= J[:, parents[1:]]
    local = np.zeros((B, NJ, 4, 4), f)
    local[:, :, :3, :3] = R
    local[:, :, :3, 3] = rel
    local[:, :, 3, 3] = 1.
    chain = [local[:, 0]]
    for i in range(1, NJ):
        chain.append(chain[parents[i]] @ local[:, i])
    G = np.stack(chain, axis=1)
    posed_joints = G[:, :, :3, 3].copy()
    A = G.copy()
    A[:, :, :3, 3] -= np.einsum('bjrc,bjc->bjr', G[:, :, :3, :3], J)
    T = np.einsum('vj,bjrc->bvrc', W, A)
    verts = np.einsum('bvrc,bvc->bvr', T[:, :, :3, :3], v_posed) + T[:, :, :3, 3]
    return verts, posed_joints


def smpl_joints(model, verts, joints24):
    """SMPL.forward's vertex joints (smplx VertexJointSelector) + the reference wrapper's extra regressor and
    joint_map (smpl_head.py:23-27) -> [B,49,3]"""
    extra_v = verts[:, model['extra_vertex_ids']]
    extra_r = np.einsum('bik,ji->bjk', verts, model['J_regressor_extra'].astype(np.float64))
    return np.concatenate([joints24, extra_v, extra_r], axis=1)[:, model['joint_map']]


def weak_perspective_to_perspective(cam, focal_length=5000., img_res=224):
    """geometry.py:447-463"""
    cam = np.asarray(cam, np.float64)
    return np.stack([cam[:, 1], cam[:, 2], 2 * focal_length / (img_res * cam[:, 0] + 1e-9)], axis=-1)


def crop_cam_to_full_img_cam(cam, bbox_height, bbox_center, img_w, img_h, focal_length, crop_res=224):
    """smplcam_head.convert_pare_to_full_img_cam (smplcam_head.py:123-139)"""
    cam = np.asarray(cam, np.float64)
    s, tx, ty = cam[:, 0], cam[:, 1], cam[:, 2]
    r = bbox_height / 224
    tz = 2 * focal_length / (r * 224 * s)
    cx = 2 * (bbox_center[:, 0] - (img_w / 2.)) / (s * bbox_height)
    cy = 2 * (bbox_center[:, 1] - (img_h / 2.)) / (s * bbox_height)
    return np.stack([tx + cx, ty + cy, tz], axis=-1)


def project(points, translation, fx, cx, cy):
    """perspective_projection with identity rotation (geometry.py:480-508 / smplcam_head.py:99-120)"""
    p = points + translation[:, None]
    p = p / p[:, :, 2:3]
    fx, cx, cy = (np.broadcast_to(np.asarray(a, np.float64), (points.shape[0],))[:, None] for a in (fx, cx, cy))
    return np.stack([fx * p[:, :, 0] + cx * p[:, :, 2], fx * p[:, :, 1] + cy * p[:, :, 2]], axis=-1)


def smpl_stage(model, rotmat, shape, cam, cliff, img_res=224, focal_length=5000., normalize_joints2d=False,
               focal=None, bbox_scale=None, bbox_center=None, img_w=None, img_h=None):
    """smpl_head.forward (smpl_head.py:45-83; cliff=False) / smplcam_head.forward (smplcam_head.py:34-96; cliff=True)"""
    f = np.float64
    verts, j24 = lbs(model, shape, rotmat)
    joints = smpl_joints(model, verts, j24)
    out = {'smpl_vertices': verts, 'smpl_joints3d': joints}
    crop_t = weak_perspective_to_perspective(cam, 5000., 224)
    out['pred_cam_t'] = crop_t
    if cliff:
        focal, bbox_scale, img_w, img_h = (np.asarray(a, f) for a in (focal, bbox_scale, img_w, img_h))
        full_t = crop_cam_to_full_img_cam(cam, bbox_scale * 200., np.asarray(bbox_center, f), img_w, img_h, focal, img_res)
        out['smpl_joints2d'] = project(joints, full_t, focal, img_w / 2., img_h / 2.)
        out['pred_fullimg_cam_t'] = full_t
    else:
        j2d = project(joints, crop_t, focal_length, 0., 0.)
        out['smpl_joints2d'] = j2d / (img_res / 2.) if normalize_joints2d else j2d
    return out
