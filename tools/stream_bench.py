"""BASELINE configs[4] shape on one GPU: a 1080p stream with D detections per frame through StreamRunner
(crop -> POCO-CLIFF/HRNet-W32 -> uncertainty post-processing -> original-image cameras), frames resident on the
device, one device -> host read of the per-frame confidence.  usage: python tools/stream_bench.py [D] [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from common import build_model  # noqa: E402
from synth import frames as C  # noqa: E402
from poco_b200 import StreamRunner  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = int(sys.argv[2]) if len(sys.argv) > 2 else 200
m = build_model('cliff_w32', 'cuda')
run = StreamRunner(m)
frames = [torch.from_numpy(C.synthetic_frame(s, 1080, 1920)).cuda() for s in range(4)]
boxes = torch.from_numpy(C.synthetic_boxes(1, D, 1080, 1920).astype(np.float32)).cuda()
for i in range(10):
    run.step(frames[i % 4], boxes)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(F):
    out = run.step(frames[i % 4], boxes)
    conf = out['confidence'].cpu()          # the per-frame result a caller reads back
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'stream: {D} detections/frame, {F} frames: {F / dt:.1f} fps ({F * D / dt:.0f} crops/s), {dt / F * 1e3:.2f} ms/frame')
