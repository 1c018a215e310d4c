"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the step right before the hot path:
the per-detection crop + normalisation of the reference demo / tester.  Nothing under poco_b200/ imports it.

Follows, line by line:
  * get_single_image_crop_demo          pocolib/utils/vibe_image_utils.py:233-267   (rot = 0, no flip)
  * generate_patch_image_cv             pocolib/utils/vibe_image_utils.py:94-108
  * gen_trans_from_patch_cv             pocolib/utils/vibe_image_utils.py:58-91
  * convert_cvimg_to_tensor / get_default_transform   vibe_image_utils.py:343-352 (ToTensor + ImageNet Normalize)
  * calculate_bbox_info, calculate_focal_length        pocolib/utils/image_utils.py:171-187
  * the batch assembly of POCOTester.run_on_image_folder pocolib/core/tester.py:181-212

The warp itself lives in a third-party dependency, OpenCV (`opencv-python`, unpinned in requirements.txt; this
image has 4.13.0): cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT) on uint8 is fixed-point arithmetic -- source
coordinates in 1/1024 pixel (AB_BITS = 10) rounded to 1/32 pixel (INTER_BITS = 5), bilinear weights
(32-fx)(32-fy)*32 ... that sum to 32768, result (sum + 16384) >> 15, taps outside the frame read 0.  That
published algorithm is restated here; parity is PINNED: tests/golden/crop_golden.npz holds the outputs of the
reference functions themselves (oracle/make_golden_crop.py, run with the reference tree and cv2) and this
restatement matches them bit for bit (tests/test_oracle_golden.py).
"""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def inverse_affine(cx, cy, bw, bh, scale, crop):
    """the 2x3 matrix cv2.warpAffine applies to destination pixels (vibe_image_utils.py:58-91 + the
    inversion inside cv::warpAffine), all in float64 except where the reference stores float32"""
    src_w, src_h = float(bw) * float(scale), float(bh) * float(scale)
    p0x, p0y = np.float32(cx), np.float32(cy)                       # src[0] = centre (stored as float32)
    p1y = np.float32(float(cy) + float(np.float32(src_h * 0.5)))    # src[1] = centre + down direction
    p2x = np.float32(float(cx) + float(np.float32(src_w * 0.5)))    # src[2] = centre + right direction
    half = float(np.float32(crop * 0.5))
    # forward map src -> dst: x' = a (x - p0x) + crop/2, y' = d (y - p0y) + crop/2
    a = half / (float(p2x) - float(p0x))
    d = half / (float(p1y) - float(p0y))
    m = [a, 0.0, half - a * float(p0x), 0.0, d, half - d * float(p0y)]
    # cv::warpAffine inverts it like this (imgwarp.cpp)
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11
    m[1] *= -D
    m[3] *= -D
    m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_u8(frame, m, crop):
    """cv2.warpAffine(frame, M, (crop, crop), INTER_LINEAR, BORDER_CONSTANT) for uint8 HWC, given the inverse M"""
    H, W = frame.shape[:2]
    xs = np.arange(crop, dtype=np.float64)
    adelta = np.rint(m[0] * xs * 1024.0).astype(np.int64)
    bdelta = np.rint(m[3] * xs * 1024.0).astype(np.int64)
    X0 = np.rint((m[1] * xs + m[2]) * 1024.0).astype(np.int64) + 16     # (indexed by the destination row y)
    Y0 = np.rint((m[4] * xs + m[5]) * 1024.0).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    x0, y0, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
    w = [(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32]
    out = np.zeros((crop, crop, 3), dtype=np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = y0 + dy, x0 + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = frame[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        v[~ok] = 0
        out += w[k][..., None] * v
    return ((out + 16384) >> 15).astype(np.uint8)


def crop_batch(frame, boxes, scale=1.2, crop=224):
    """frame uint8 [H, W, 3] RGB, boxes [[cx, cy, w, h], ...] -> the batch dict of tester.py:205-212 (numpy)"""
    H, W = frame.shape[:2]
    imgs, info, focal, scales, centers, shapes = [], [], [], [], [], []
    for cx, cy, bw, bh in np.asarray(boxes, dtype=np.float64):
        raw = warp_u8(frame, inverse_affine(cx, cy, bw, bh, scale, crop), crop)
        t = raw.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)               # ToTensor
        imgs.append((t - MEAN[:, None, None]) / STD[:, None, None])                      # Normalize
        s = max(bw, bh) / 200.0
        f = float((W ** 2 + H ** 2) ** 0.5)
        b = s * 200
        bi = np.array([cx - W / 2.0, cy - H / 2.0, b])
        bi[:2] = bi[:2] / f * 2.8
        bi[2] = (bi[2] - 0.24 * f) / (0.06 * f)
        info.append(bi.astype(np.float32))
        focal.append(f)
        scales.append(s)
        centers.append([cx, cy])
        shapes.append([H, W])
    return {'img': np.stack(imgs).astype(np.float32), 'bbox_info': np.stack(info),
            'focal_length': np.asarray(focal, dtype=np.float32), 'scale': np.asarray(scales, dtype=np.float32),
            'center': np.asarray(centers, dtype=np.float32), 'orig_shape': np.asarray(shapes, dtype=np.float32)}


from synth.frames import synthetic_boxes, synthetic_frame  # noqa: E402,F401  (data generators, kept importable from here)
