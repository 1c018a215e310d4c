"""run ONE conv shape a few times through the C ABI (for ncu captures).
usage: python tools/one_conv.py cin cout k stride H res [B=256] [reps=3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402

cin, cout, k, st, H, res = (int(v) for v in sys.argv[1:7])
B = int(sys.argv[7]) if len(sys.argv) > 7 else 256
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
s = torch.cuda.current_stream().cuda_stream
Ho = (H + 2 * (k // 2) - k) // st + 1
a = engine.alloc_act(cin, B, H, H, 'cuda')
engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
o = engine.alloc_act(cout, B, Ho, Ho, 'cuda')
r = engine.alloc_act(cout, B, Ho, Ho, 'cuda') if res else None
w = (torch.randn(k * k, cin // 8, cout, 8, device='cuda') * 0.05).half()
b = torch.zeros(cout, device='cuda')
d = L.Conv(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), r.ptr if r else None, r.plane_stride if r else 0,
           k, k, st, k // 2, 1, 0, 0, 0)
op = L.make_op(d)
for _ in range(reps):
    L.run_op(op, s)
torch.cuda.synchronize()
print('done')
