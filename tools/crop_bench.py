"""throughput of the GPU crop + normalise op: 256 detections on a 1080p frame, CUDA events"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from synth import frames as C  # noqa: E402
from poco_b200 import crop_batch  # noqa: E402

fr = torch.from_numpy(C.synthetic_frame(5, 1080, 1920)).cuda()
bx = torch.from_numpy(C.synthetic_boxes(7, 256, 1080, 1920).astype(np.float32)).cuda()
for _ in range(3):
    crop_batch(fr, bx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = crop_batch(fr, bx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
by = out['img'].numel() * 4
print(f'crop_batch 256 crops from 1080p: {ms * 1e3:.1f} us  ({256 / ms * 1e3:.0f} crops/s, {by / ms / 1e6:.0f} GB/s written)')
