"""GPU crop + normalise: the caller-side step right before POCO.forward.

`crop_batch(frame, boxes)` replaces the per-detection CPU loop of the reference demo / tester
(pocolib/core/tester.py:181-212: get_single_image_crop_demo + calculate_bbox_info + calculate_focal_length) for
a frame that already lives in device memory and returns the batch dict POCO.forward takes.  The crops are
bit-identical to the reference's (cv2.warpAffine's fixed-point bilinear is reproduced exactly by the kernel,
poco_b200/csrc/preprocess.cu); there is no CPU fallback.
"""
import torch

from . import _lib as L


def crop_batch(frame, boxes, scale=1.2, crop=224, stream=None):
    """frame: uint8 CUDA tensor [H, W, 3] (RGB); boxes: [n, 4] = (cx, cy, w, h) in pixels.
    -> {'img' [n,3,crop,crop], 'bbox_info' [n,3], 'focal_length' [n], 'scale' [n], 'center' [n,2], 'orig_shape' [n,2]}"""
    if not (isinstance(frame, torch.Tensor) and frame.is_cuda and frame.dtype == torch.uint8 and frame.dim() == 3
            and frame.shape[2] == 3):
        raise L.PocoError('crop_batch needs a uint8 CUDA frame of shape [H, W, 3] (poco_b200 has no CPU path)')
    frame = frame.contiguous()
    boxes = torch.as_tensor(boxes, dtype=torch.float32).to(frame.device).contiguous().view(-1, 4)
    n = boxes.shape[0]
    if n == 0:
        raise ValueError('crop_batch: no detections')
    dev = frame.device
    out = {'img': torch.empty(n, 3, crop, crop, dtype=torch.float32, device=dev),
           'bbox_info': torch.empty(n, 3, dtype=torch.float32, device=dev),
           'focal_length': torch.empty(n, dtype=torch.float32, device=dev),
           'scale': torch.empty(n, dtype=torch.float32, device=dev),
           'center': torch.empty(n, 2, dtype=torch.float32, device=dev),
           'orig_shape': torch.empty(n, 2, dtype=torch.float32, device=dev)}
    d = L.Crop(frame.data_ptr(), frame.shape[0], frame.shape[1], boxes.data_ptr(), n, crop, float(scale), 0,
               out['img'].data_ptr(), out['bbox_info'].data_ptr(), out['focal_length'].data_ptr(),
               out['scale'].data_ptr(), out['center'].data_ptr(), out['orig_shape'].data_ptr())
    with torch.cuda.device(dev):
        L.run_op(d, stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    return out


def uncert_post(var_pose, backbone, kinematic=False, return_conf=False, sensitivity_threshold=0.40, stream=None):
    """POCOUtils.prepare_uncert + get_global_uncert (pocolib/utils/poco_utils.py:50-94, tester.py:243-245) on the
    device.  var_pose: f32 CUDA [n, 24] from POCO.forward; backbone: the model's '<backbone>-<head>' string.
    -> (prepared [n,24], thresholded [n,24], global_var [n]), all on the device: no host synchronisation."""
    if not (isinstance(var_pose, torch.Tensor) and var_pose.is_cuda and var_pose.dtype == torch.float32
            and var_pose.dim() == 2 and var_pose.shape[1] == 24):
        raise L.PocoError('uncert_post needs an f32 CUDA tensor of shape [n, 24] (poco_b200 has no CPU path)')
    v = var_pose.contiguous()
    n = v.shape[0]
    prepared, thr = torch.empty_like(v), torch.empty_like(v)
    glob = torch.empty(n, dtype=torch.float32, device=v.device)
    d = L.UncertPost(v.data_ptr(), n, 1 if 'cliff' in backbone else 0, int(bool(kinematic)), int(bool(return_conf)),
                     float(sensitivity_threshold), 0, prepared.data_ptr(), thr.data_ptr(), glob.data_ptr())
    with torch.cuda.device(v.device):
        L.run_op(d, stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    return prepared, thr, glob
