"""phase timeline of CTA 0 over back-to-back launches of one conv (POCO_CONV_DEBUG bit 64): where the fixed per-launch cost goes.
usage: python tools/conv_timeline.py cin cout k stride H res B [launches=4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402

cin, cout, k, st, H, res, B = (int(v) for v in sys.argv[1:8])
n = int(sys.argv[8]) if len(sys.argv) > 8 else 4
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
prof = torch.zeros(128, dtype=torch.int64, device='cuda')
os.environ['POCO_CONV_PROF'] = str(prof.data_ptr())
os.environ['POCO_CONV_DEBUG'] = '0'
for _ in range(3):
    L.run_op(op, s)
torch.cuda.synchronize()
os.environ['POCO_CONV_DEBUG'] = '64'
for use_graph in (0, 1):
    prof.zero_()
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                L.run_op(op, s if not use_graph else torch.cuda.current_stream().cuda_stream)
        prof.zero_()
        g.replay()
    else:
        for _ in range(n):
            L.run_op(op, s)
    torch.cuda.synchronize()
    v = prof.cpu().tolist()
    cnt = v[127]
    rows = sorted([v[16 + i * 8:24 + i * 8] for i in range(min(cnt, 12))])
    print(f'# {cin}->{cout} k{k} s{st} h{H} res{res} B{B} graph={use_graph}: per launch, us relative to kernel entry; gap = entry - previous exit')
    print('launch,gap_us,setup_done,first_data,first_unit_issued,first_acc_ready,first_store,last_store,exit')
    prev = None
    for i, st_ in enumerate(rows):
        t0 = st_[0]
        rel = [(x - t0) / 1e3 for x in st_]
        gap = (t0 - prev) / 1e3 if prev else 0.0
        print(f'{i},{gap:.2f},' + ','.join(f'{x:.2f}' for x in rel[1:]))
        prev = st_[7]
