"""time the stride-2 3x3 convs of HRNet-W32 at batch 256 (POCO_B200_S2=0 -> generic gather producer)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from gpu_util import sync_or_die  # noqa: E402
from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402

B = 256
CASES = [(16, 64, 224, False), (64, 64, 112, False), (32, 32, 56, False), (32, 64, 56, True), (64, 128, 28, True),
         (32, 128, 28, True), (256, 64, 56, False), (32, 32, 28, False), (64, 64, 28, False)]
s = torch.cuda.current_stream().cuda_stream
print('case,us,TFLOPs,GBs', flush=True)
for cin, cout, H, res in CASES:
    Ho = H // 2
    a = engine.alloc_act(cin, B, H, H, 'cuda')
    engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
    o = engine.alloc_act(cout, B, Ho, Ho, 'cuda')
    r = engine.alloc_act(cout, B, Ho, Ho, 'cuda') if res else None
    w = (torch.randn(9, cin // 8, cout, 8, device='cuda') * 0.05).half()
    b = torch.zeros(cout, device='cuda')
    d = L.Conv(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), r.ptr if r else None, r.plane_stride if r else 0,
               3, 3, 2, 1, 1, 0)
    op = L.make_op(d)
    for _ in range(3):
        L.run_op(op, s)
    sync_or_die(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.run_op(op, s)
    e1.record()
    sync_or_die(20)
    us = e0.elapsed_time(e1) * 100
    fl = 2.0 * B * Ho * Ho * cout * cin * 9
    by = 2.0 * B * (H * H * cin + Ho * Ho * cout * (2 if res else 1))
    print(f'{cin}->{cout} s2 h{H} res{int(res)},{us:.1f},{fl / us / 1e6:.1f},{by / us / 1e3:.0f}', flush=True)
