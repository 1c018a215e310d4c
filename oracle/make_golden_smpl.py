"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/smpl_cam_golden.npz from the REFERENCE's own camera and
projection functions (the part of the SMPL mesh stage that lives in the reference tree):
    python -m oracle.make_golden_smpl
  pocolib.utils.geometry.convert_weak_perspective_to_perspective / perspective_projection (geometry.py:447-508)
  pocolib.models.head.smplcam_head.convert_pare_to_full_img_cam / perspective_projection (smplcam_head.py:99-139)
called the way smpl_head.forward (:64-83) and smplcam_head.forward (:58-94) call them, on seeded joints.  The LBS
itself (smplx) cannot be run here -- see oracle/smpl_oracle.py."""
import os
import sys

import numpy as np
import torch

from . import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'smpl_cam_golden.npz')


def main():
    root = R.find_reference_root()
    assert root, 'reference tree not found'
    R._install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    import pocolib.utils.geometry as G
    importlib.import_module('pocolib.models.head')
    SC = sys.modules['pocolib.models.head.smplcam_head']      # (the package re-exports a class of the same name)
    r = np.random.default_rng(11)
    n = 16
    joints = torch.from_numpy((r.standard_normal((n, 49, 3)) * np.array([0.3, 0.5, 0.15])).astype(np.float32))
    cam = torch.from_numpy(np.stack([r.uniform(0.5, 1.3, n), r.normal(0, 0.1, n), r.normal(0, 0.1, n)], -1).astype(np.float32))
    out = {'joints': joints.numpy(), 'cam': cam.numpy()}
    # smpl_head.forward :64-83 (PARE)
    cam_t = G.convert_weak_perspective_to_perspective(cam)
    j2d = G.perspective_projection(joints, rotation=torch.eye(3).unsqueeze(0).expand(n, -1, -1), translation=cam_t,
                                   focal_length=5000., camera_center=torch.zeros(n, 2))
    out['pare_cam_t'] = cam_t.numpy()
    out['pare_joints2d'] = j2d.numpy()
    out['pare_joints2d_norm'] = (j2d / (224 / 2.)).numpy()
    # smplcam_head.forward :58-94 (CLIFF); intrinsics built on the CPU here (the reference hard-codes .cuda())
    img_w = torch.from_numpy(r.choice([640., 1280., 1920.], n).astype(np.float32))
    img_h = torch.from_numpy(r.choice([480., 720., 1080.], n).astype(np.float32))
    center = torch.from_numpy(np.stack([r.uniform(100, 500, n), r.uniform(100, 400, n)], -1).astype(np.float32))
    scale = torch.from_numpy(r.uniform(0.6, 3.0, n).astype(np.float32))
    focal = torch.sqrt(img_w ** 2 + img_h ** 2)                      # calculate_focal_length (image_utils.py:171-172)
    K = torch.eye(3).repeat(n, 1, 1).float()
    K[:, 0, 0] = focal
    K[:, 1, 1] = focal
    K[:, 0, 2] = img_w / 2.
    K[:, 1, 2] = img_h / 2.
    full_t = SC.convert_pare_to_full_img_cam(pare_cam=cam.detach().clone(), bbox_height=scale * 200., bbox_center=center,
                                             img_w=img_w, img_h=img_h, focal_length=K[:, 0, 0], crop_res=224)
    j2d = SC.perspective_projection(joints, rotation=torch.eye(3).unsqueeze(0).expand(n, -1, -1), translation=full_t,
                                    cam_intrinsics=K)
    out.update(img_w=img_w.numpy(), img_h=img_h.numpy(), center=center.numpy(), scale=scale.numpy(), focal=focal.numpy(),
               cliff_full_t=full_t.numpy(), cliff_joints2d=j2d.numpy(),
               cliff_cam_t=G.convert_weak_perspective_to_perspective(cam).numpy())
    import pocolib.core.constants as K
    out['joint_map'] = np.array([K.JOINT_MAP[n] for n in K.JOINT_NAMES], np.int32)      # smpl_head.py:17 (constants.py:15-93)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
