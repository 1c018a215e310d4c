"""bring-up of the conv chain kernel: a ladder of cases with flushed progress output"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from gpu_util import conv_reference, run_conv, stream, sync_or_die  # noqa: E402
from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402
import test_gpu_ops as T  # noqa: E402


def say(*a):
    print(*a, flush=True)


say('single conv 32->32 res')
x, w, b, r = T._case_tensors(32, 32, 3, 1, 56, 2, True)
out = run_conv(x, w, b, 1, None, 1, r)
say('  err', float((out - conv_reference(x, w, b, 1, None, 1, r)).abs().max()))

for ch, H, N, mc in [(32, 56, 12, 0), (64, 28, 20, 0), (128, 14, 40, 0), (256, 7, 64, 0), (32, 56, 5, 7)]:
    outs = []
    for chained in (False, True):
        bld, xin, xout, x0, sd = T._branch_ops(ch, H, N, chained, mc)
        say(f'case ch{ch} H{H} N{N} mc{mc} chained={chained} ops={len(bld.ops)}')
        engine.act_view(xin)[:, :, 1:H + 1, 1:H + 1, :] = x0.cuda().half().view(N, ch // 8, 8, H, H).permute(1, 0, 3, 4, 2)
        for op in bld.ops:
            L.run_op(op, stream())
        try:
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001
            say('  CUDA error:', repr(e)[:300])
            sys.exit(1)
        outs.append(engine.from_planar(xout).cpu())
        say('  done, out absmax', float(outs[-1].abs().max()))
    say('  chained == separate:', bool(torch.equal(outs[0], outs[1])), 'max diff', float((outs[0] - outs[1]).abs().max()))
