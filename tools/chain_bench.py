"""time an HRNet branch (4 BasicBlocks = 8 convs) as 8 launches vs one chained launch, at several batch sizes
usage: python tools/chain_bench.py [max_ctas]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from gpu_util import stream, sync_or_die  # noqa: E402
from poco_b200 import _lib as L  # noqa: E402
import test_gpu_ops as T  # noqa: E402

mc = int(sys.argv[1]) if len(sys.argv) > 1 else 0
GROUPS = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1]
print('ch,H,B,max_ctas,groups,separate_us_per_conv,chain_us_per_conv', flush=True)
for ch, H in [(32, 56), (64, 28), (128, 14), (256, 7)]:
    for B, groups in [(16, 1), (64, 1)] + [(256, g) for g in GROUPS]:
        res = []
        for chained in (False, True):
            bld, xin, xout, x0, sd = T._branch_ops(ch, H, B, chained, mc, groups=groups)
            for _ in range(3):
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
            res.append(e0.elapsed_time(e1) * 1e3 / 5 / 8)
            del bld
        print(f'{ch},{H},{B},{mc},{groups},{res[0]:.1f},{res[1]:.1f}', flush=True)
