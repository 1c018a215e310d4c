"""cycle accounting of the two MMA issuer warps of one conv launch (POCO_CONV_DEBUG bit 32).
usage: python tools/conv_prof.py cin cout k stride H res B "debug,list" """
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402

cin, cout, k, st, H, res, B = (int(v) for v in sys.argv[1:8])
DBG = [int(x) for x in sys.argv[8].split(',')]
s = torch.cuda.current_stream().cuda_stream
Ho = (H + 2 * (k // 2) - k) // st + 1
SPLIT = os.environ.get('PROF_SPLIT') == '1'
a = engine.alloc_act(cin, B, H, H, 'cuda', SPLIT)
engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
o = engine.alloc_act(cout, B, Ho, Ho, 'cuda', SPLIT)
r = engine.alloc_act(cout, B, Ho, Ho, 'cuda', SPLIT) if res else None
w = (torch.randn((2 if SPLIT else 1) * k * k, cin // 8, cout, 8, device='cuda') * 0.05).half()
b = torch.zeros(cout, device='cuda')
d = L.Conv(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), r.ptr if r else None, r.plane_stride if r else 0,
           k, k, st, k // 2, 1, 0, 0, 0, r.ptr_lo if r else None)
op = L.make_op(d)
prof = torch.zeros(16, dtype=torch.int64, device='cuda')
os.environ['POCO_CONV_PROF'] = str(prof.data_ptr())
print('case,debug,issuer,ctas,units_per_cta,total_cyc_per_cta,per_unit: acc_wait,full_wait,issue,other')
for dbg in DBG:
    os.environ['POCO_CONV_DEBUG'] = str(dbg)
    L.run_op(op, s)
    torch.cuda.synchronize()
    prof.zero_()
    os.environ['POCO_CONV_DEBUG'] = str(dbg | 32)
    L.run_op(op, s)
    torch.cuda.synchronize()
    v = prof.cpu().tolist()
    for mw in range(2):
        tot, acc, full, issue, units, ctas = v[mw * 8:mw * 8 + 6]
        if ctas == 0:
            continue
        u = max(units, 1)
        print(f'{cin}->{cout} k{k} s{st} h{H} res{res} B{B},{dbg},{mw},{ctas},{units / ctas:.1f},{tot / ctas:.0f},'
              f'{acc / u:.0f},{full / u:.0f},{issue / u:.0f},{(tot - acc - full - issue) / u:.0f}', flush=True)
