"""launch a few representative convs once each (for ncu captures): python tools/prof_cases.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
# (cin, cout, k, stride, H, res)
CASES = [(32, 32, 3, 1, 56, True), (64, 64, 3, 1, 28, True), (128, 128, 3, 1, 14, True), (256, 256, 3, 1, 7, True),
         (32, 64, 3, 2, 56, True), (64, 256, 1, 1, 56, True), (256, 256, 3, 1, 56, False)]
s = torch.cuda.current_stream().cuda_stream
for cin, cout, k, st, H, res in CASES:
    Ho = (H + 2 * (k // 2) - k) // st + 1
    a = engine.alloc_act(cin, B, H, H, 'cuda')
    engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
    o = engine.alloc_act(cout, B, Ho, Ho, 'cuda')
    r = engine.alloc_act(cout, B, Ho, Ho, 'cuda') if res else None
    w = (torch.randn(k * k, cin // 8, cout, 8, device='cuda') * 0.05).half()
    b = torch.zeros(cout, device='cuda')
    d = L.Conv(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), r.ptr if r else None, r.plane_stride if r else 0,
               k, k, st, k // 2, 1, 0)
    op = L.make_op(d)
    for _ in range(2):
        L.run_op(op, s)
    torch.cuda.synchronize()
    print('ran', cin, cout, k, st, H, flush=True)
