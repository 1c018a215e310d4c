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
    (uint8 [H, W, 3]) already in device memory, `boxes` the detections [n, 4] = (cx, cy, w, h).

    graph=True (default): the whole per-frame body -- crop kernel, the plan's ~360 launches, uncertainty
    post-processing, camera conversion (and the mesh stage when the model has one on the device) -- is captured ONCE
    per (frame size, detection-count bucket) into a CUDA graph over static input buffers; a step is then one frame copy
    into the static buffer (6 MB at HBM speed), one boxes copy and one graph launch, which is what keeps small-batch
    frames (a few detections: launch-latency bound, not compute bound) from paying ~360 host-side launches.  Detection
    counts are padded to POCO.bucket(n) with copies of the first box; the padding rows are dropped.  The returned
    tensors are views into the graph's static outputs: valid until the next step() of the same bucket (clone to keep)."""

    def __init__(self, model, bbox_scale=1.2, crop=224, kinematic_uncert=False, sensitivity_threshold=0.40,
                 clip_global=True, graph=True):
        """clip_global: np.clip(variance_global, 0, 0.99) as the image-folder loop does (tester.py:245); the tracked-video
        loop (tester.py:418-421) keeps the raw value -- pass False for that behaviour."""
        self.clip_global = bool(clip_global)
        self.model = model.eval()
        self.bbox_scale, self.crop = float(bbox_scale), int(crop)
        self.kinematic, self.threshold = bool(kinematic_uncert), float(sensitivity_threshold)
        self.backbone = f'{model.backbone_name}-{model.head_name}'
        self.use_graph = bool(graph)
        self._graphs = {}           # (H, W, padded n, device) -> (graph, static frame, static boxes, static outputs)
        self.max_graphs = 8

    @torch.no_grad()
    def _body(self, frame, boxes):
        H, W = int(frame.shape[0]), int(frame.shape[1])
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

    @torch.no_grad()
    def step(self, frame, boxes):
        boxes = torch.as_tensor(boxes, dtype=torch.float32).to(frame.device).view(-1, 4)
        if not self.use_graph:
            return self._body(frame, boxes)
        n = int(boxes.shape[0])
        if n == 0:
            raise ValueError('StreamRunner.step: no detections')
        npad = self.model.bucket(n)
        key = (int(frame.shape[0]), int(frame.shape[1]), npad, str(frame.device))
        ent = self._graphs.pop(key, None)
        if ent is None:
            while len(self._graphs) >= self.max_graphs:
                del self._graphs[next(iter(self._graphs))]      # least recently used
            s_frame = torch.empty_like(frame)
            s_boxes = torch.empty(npad, 4, dtype=torch.float32, device=frame.device)
            s_frame.copy_(frame)
            s_boxes[:n].copy_(boxes)
            s_boxes[n:] = boxes[0]
            for _ in range(2):                                  # builds the plan, one-time attribute set-up
                self._body(s_frame, s_boxes)
            torch.cuda.synchronize(frame.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s_out = self._body(s_frame, s_boxes)
            ent = (g, s_frame, s_boxes, s_out)
        self._graphs[key] = ent
        g, s_frame, s_boxes, s_out = ent
        s_frame.copy_(frame)
        s_boxes[:n].copy_(boxes)
        if npad > n:
            s_boxes[n:] = boxes[0]
        g.replay()
        return {k: (v[:n] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == npad else v) for k, v in s_out.items()}
