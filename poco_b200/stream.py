"""Video-stream driver (BASELINE configs[4]): frame + detections -> crops -> POCO -> confidence and
original-image cameras, one stream per GPU, nothing leaves the device until the caller asks.

It is the caller side of the hot path rebuilt on the three device ops of this package -- `crop_batch` (f1),
`POCO.forward` (a1-a12) and `uncert_post` (f3) -- and mirrors the per-frame body of the reference loops
(pocolib/core/tester.py:181-245 for image folders, :396-455 for tracked videos).  Detector and tracker are
third-party (requirements.txt:29-30, unpinned) and stay outside: boxes are an input.
"""
import torch

from .preprocess import crop_batch, uncert_post


def convert_crop_cam_to_orig_img(cam, bbox, img_width, img_height):
    """weak-perspective camera of the crop -> camera of the original image (torch restatement of
    pocolib/utils/demo_utils.py:249-266: cam [n,3] = (s, tx, ty), bbox [n,>=3] = (cx, cy, h, ...))"""
    cx, cy, h = bbox[:, 0], bbox[:, 1], bbox[:, 2]
    hw, hh = img_width / 2., img_height / 2.
    sx = cam[:, 0] * (1. / (img_width / h))
    sy = cam[:, 0] * (1. / (img_height / h))
    tx = ((cx - hw) / hw / sx) + cam[:, 1]
    ty = ((cy - hh) / hh / sy) + cam[:, 2]
    return torch.stack([sx, sy, tx, ty]).T


class StreamRunner:
    """one video stream on one GPU.  step(frame, boxes) returns device tensors; `frame` is the decoded RGB frame
    (uint8 [H, W, 3]) already in device memory, `boxes` the detections [n, 4] = (cx, cy, w, h)."""

    def __init__(self, model, bbox_scale=1.2, crop=224, kinematic_uncert=False, sensitivity_threshold=0.40,
                 clip_global=True):
        """clip_global: np.clip(variance_global, 0, 0.99) as the image-folder loop does (tester.py:245); the tracked-video
        loop (tester.py:418-421) keeps the raw value -- pass False for that behaviour."""
        self.clip_global = bool(clip_global)
        self.model = model.eval()
        self.bbox_scale, self.crop = float(bbox_scale), int(crop)
        self.kinematic, self.threshold = bool(kinematic_uncert), float(sensitivity_threshold)
        self.backbone = f'{model.backbone_name}-{model.head_name}'

    @torch.no_grad()
    def step(self, frame, boxes):
        H, W = int(frame.shape[0]), int(frame.shape[1])
        boxes = torch.as_tensor(boxes, dtype=torch.float32).to(frame.device).view(-1, 4)
        batch = crop_batch(frame, boxes, scale=self.bbox_scale, crop=self.crop)
        out = self.model(batch)
        var, _, var_global = uncert_post(out['var_pose'], self.backbone, kinematic=self.kinematic,
                                         sensitivity_threshold=self.threshold)
        out['variance'] = var                                   # tester.py:243
        if self.clip_global:
            var_global = torch.clamp(var_global, 0.0, 0.99)     # tester.py:245
        out['variance_global'] = var_global                     # tester.py:244
        out['confidence'] = 1.0 - var_global
        out['orig_cam'] = convert_crop_cam_to_orig_img(out['pred_cam'], boxes, W, H)      # tester.py:216-221
        return out
