"""throughput of the device SMPL mesh stage: poco_smpl_run on 256 crops, CUDA events on the launching
stream, per-kernel split; one JSON line.  python tools/smpl_bench.py [n] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from synth import smpl_model as O  # noqa: E402
from poco_b200 import smpl as S  # noqa: E402
from test_smpl import inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
m = O.synthetic_model(0)
st = S.DeviceSmplStage('cliff', m).to('cuda')
d = inputs(n, seed=1, cliff=True)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
args = (t(d['rotmat']), t(d['shape']), t(d['cam']))
kw = dict(focal_length=t(d['focal']), bbox_scale=t(d['bbox_scale']), bbox_center=t(d['bbox_center']), img_w=t(d['img_w']),
          img_h=t(d['img_h']))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')    # > 126 MB L2
for _ in range(3):
    st(*args, **kw)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    flush.zero_()                                                   # cold L2 between timed iterations
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = st(*args, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = sorted(ms)
med = ms[len(ms) // 2]
out_bytes = sum(v.numel() * 4 for v in out.values())
model_bytes = sum(getattr(st, 'm_' + k).numel() * 4 for k in st._names)
flops = 2.0 * n * st.vp * (217 * 3 + 24 * 12 + 9)
print(json.dumps({'op': 'poco_smpl_run', 'crops': n, 'median_ms': round(med, 4), 'best_ms': round(ms[0], 4),
                  'crops_per_s': round(n / med * 1e3), 'launches': 3,
                  'algorithmic_bytes': out_bytes + model_bytes + n * (216 + 10 + 3 + 7) * 4,
                  'achieved_GBs': round((out_bytes + model_bytes) / med / 1e6, 1),
                  'fp32_gflop': round(flops / 1e9, 3), 'achieved_fp32_TFLOPs': round(flops / med / 1e9, 2),
                  'l2': 'flushed between iterations (256 MB memset)', 'timing': 'CUDA events incl. output allocation (caching allocator)'}))
