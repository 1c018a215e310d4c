"""cycle accounting of the MMA issuers (POCO_CONV_DEBUG bit 32) for an HRNet branch run as 8 launches vs one chain.
usage: python tools/chain_prof.py ch H B"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from gpu_util import stream, sync_or_die  # noqa: E402
from poco_b200 import _lib as L  # noqa: E402
import test_gpu_ops as T  # noqa: E402

ch, H, B = (int(v) for v in sys.argv[1:4])
prof = torch.zeros(128, dtype=torch.int64, device='cuda')
os.environ['POCO_CONV_PROF'] = str(prof.data_ptr())
print('mode,issuer,ctas,units_per_cta,total_cyc_per_cta,per_unit: acc_wait,full_wait,issue,other,  us_per_conv')
for chained in (False, True):
    bld, xin, xout, x0, sd = T._branch_ops(ch, H, B, chained, 0)
    os.environ['POCO_CONV_DEBUG'] = '0'
    for _ in range(2):
        for op in bld.ops:
            L.run_op(op, stream())
    sync_or_die(30)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for op in bld.ops:
            L.run_op(op, stream())
    e1.record()
    sync_or_die(30)
    us = e0.elapsed_time(e1) * 1e3 / 5 / 8
    prof.zero_()
    os.environ['POCO_CONV_DEBUG'] = '32'
    for op in bld.ops:
        L.run_op(op, stream())
    sync_or_die(30)
    v = prof.cpu().tolist()
    for mw in range(4):
        tot, acc, full, issue, units, ctas = v[mw * 8:mw * 8 + 6]
        if ctas == 0:
            continue
        u = max(units, 1)
        print(f'{"chain" if chained else "separate"},{mw},{ctas},{units / ctas:.1f},{tot / ctas:.0f},'
              f'{acc / u:.0f},{full / u:.0f},{issue / u:.0f},{(tot - acc - full - issue) / u:.0f},  {us:.1f}', flush=True)
    os.environ['POCO_CONV_DEBUG'] = '0'
    del bld
