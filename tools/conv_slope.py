"""marginal cost per tile vs fixed launch cost: time one conv shape at several batch sizes and role-skipping masks.
usage: python tools/conv_slope.py <tag> cin cout k stride H res "debug,list" "batch,list" """
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from gpu_util import sync_or_die  # noqa: E402

tag = sys.argv[1]
cin, cout, k, st, H, res = (int(v) for v in sys.argv[2:8])
DBG = [int(x) for x in sys.argv[8].split(',')]
BS = [int(x) for x in sys.argv[9].split(',')]
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in BS:
    Ho = (H + 2 * (k // 2) - k) // st + 1
    a = engine.alloc_act(cin, B, H, H, 'cuda')
    if os.environ.get('SLOPE_ZERO') != '1':
        engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
    o = engine.alloc_act(cout, B, Ho, Ho, 'cuda')
    r = engine.alloc_act(cout, B, Ho, Ho, 'cuda') if res else None
    w = (torch.randn(k * k, cin // 8, cout, 8, device='cuda') * (0.0 if os.environ.get('SLOPE_ZERO') == '1' else 0.05)).half()
    b = torch.zeros(cout, device='cuda')
    d = L.Conv(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), r.ptr if r else None, r.plane_stride if r else 0,
               k, k, st, k // 2, 1, 0, 0, 0)
    op = L.make_op(d)
    for dbg in DBG:
        os.environ['POCO_CONV_DEBUG'] = str(dbg)
        for _ in range(3):
            L.run_op(op, s)
        sync_or_die(20)
        best = 1e9
        for _ in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.run_op(op, s)
            e1.record()
            sync_or_die(20)
            best = min(best, e0.elapsed_time(e1))
        print(f'{tag},{cin}->{cout} k{k} s{st} h{H} res{res},{B},{dbg},{best * 1e3:.1f}', flush=True)
    os.environ['POCO_CONV_DEBUG'] = '0'
