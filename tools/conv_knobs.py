"""time the BasicBlock conv shapes through the C ABI under the current environment knobs (one process per knob set:
the library caches its environment reads).  usage: python tools/conv_knobs.py <tag> [debug list]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gpu_util import sync_or_die  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'default'
DBG = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0]
B = 256
CASES = [(32, 32, 3, 1, 56, False), (32, 32, 3, 1, 56, True), (64, 64, 3, 1, 28, False), (64, 64, 3, 1, 28, True),
         (128, 128, 3, 1, 14, True), (256, 256, 3, 1, 7, True), (64, 256, 1, 1, 56, True), (256, 64, 1, 1, 56, False),
         (256, 256, 3, 1, 56, False), (32, 64, 3, 2, 56, True), (64, 64, 3, 2, 112, False)]
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for cin, cout, k, st, H, res in CASES:
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
    for dbg in DBG:
        os.environ['POCO_CONV_DEBUG'] = str(dbg)
        for _ in range(3):
            L.run_op(op, s)
        sync_or_die(20)
        tot = 0.0
        for _ in range(10):         # cold L2 for every timed launch
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.run_op(op, s)
            e1.record()
            sync_or_die(20)
            tot += e0.elapsed_time(e1)
        us = tot * 100
        print(f'{tag},{cin}->{cout} k{k} s{st} h{H} res{int(res)},{dbg},{us:.1f}', flush=True)
    os.environ['POCO_CONV_DEBUG'] = '0'
