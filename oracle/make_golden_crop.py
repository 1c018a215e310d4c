"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/crop_golden.npz from the REFERENCE's own crop functions
(needs the reference tree and cv2; run in the build container):   python -m oracle.make_golden_crop

For every synthetic detection it calls pocolib.utils.vibe_image_utils.get_single_image_crop_demo and
pocolib.utils.image_utils.calculate_bbox_info / calculate_focal_length exactly as
pocolib/core/tester.py:181-212 does, and stores inputs + outputs."""
import importlib
import os
import sys
import types

import numpy as np

from . import crop_oracle as C
from . import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'crop_golden.npz')


def import_reference_utils():
    root = R.find_reference_root()
    assert root, 'reference tree not found'
    R._install_shims()
    for n in ('skimage', 'skimage.util', 'skimage.util.shape', 'skimage.transform', 'trimesh', 'trimesh.visual', 'jpeg4py',
              'matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.gridspec', 'scipy.misc'):
        if n not in sys.modules:
            try:
                importlib.import_module(n)
            except Exception:       # noqa: BLE001 -- import-only dependencies of the two utility modules
                sys.modules[n] = types.ModuleType(n)
    sys.modules['skimage.util.shape'].__dict__.setdefault('view_as_windows', None)
    sys.modules['skimage.transform'].__dict__.setdefault('rotate', None)
    sys.modules['skimage.transform'].__dict__.setdefault('resize', None)
    sys.modules['trimesh.visual'].__dict__.setdefault('color', None)
    if root not in sys.path:
        sys.path.insert(0, root)
    import pocolib.utils.image_utils as I
    import pocolib.utils.vibe_image_utils as V
    return V, I


def main():
    V, I = import_reference_utils()
    frame = C.synthetic_frame(0)
    boxes = C.synthetic_boxes(0)
    H, W = frame.shape[:2]
    scale = 1.2
    imgs, info, focal = [], [], []
    for cx, cy, bw, bh in boxes:
        bbox = [float(cx), float(cy), float(bw), float(bh)]
        norm_img, raw_img, _ = V.get_single_image_crop_demo(frame, bbox, kp_2d=None, scale=scale, crop_size=224)
        imgs.append(norm_img.float().numpy())
        s = max(bbox[2], bbox[3]) / 200.
        info.append(I.calculate_bbox_info([bbox[0], bbox[1]], s, [H, W]))
        focal.append(I.calculate_focal_length(H, W))
    np.savez_compressed(OUT, frame=frame, boxes=boxes.astype(np.float32), scale=np.float32(scale),
                        img=np.stack(imgs).astype(np.float32), bbox_info=np.stack(info).astype(np.float32),
                        focal_length=np.asarray(focal, dtype=np.float32))
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
