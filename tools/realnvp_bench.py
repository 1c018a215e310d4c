"""RealNVP log_prob for 256 crops x 24 joints (6144 rows), conditional flow with a 1024-wide context: the kernel walking
the full [64][9+1024] first layers per 8-row CTA against the context hoist (one GEMM per call + the 9-column part).
usage: python tools/realnvp_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from common import build_model  # noqa: E402
from poco_b200 import _lib as L  # noqa: E402

m = build_model('cliff_w32', 'cuda')
B, J = 256, 24
g = torch.Generator().manual_seed(0)
x = torch.randn(B * J, 9, generator=g).cuda()
ctx = torch.randn(B, m.context_dim, generator=g).cuda()
ctx_rep = torch.repeat_interleave(ctx, J, 0).contiguous()
params, nl = m._flow_params(x.device)[:2]
s = torch.cuda.current_stream().cuda_stream
out = torch.empty(B * J, device='cuda')


def legacy():
    L.run_op(L.RealNVP(x.data_ptr(), ctx_rep.data_ptr(), params.data_ptr(), out.data_ptr(), None, None,
                       B * J, 9, ctx.shape[1], 64, nl, 0, None, 1, 0), s)


def hoisted():
    return m.flow_log_prob(x, ctx, rows_per_ctx=J)


res = {}
for name, fn in (('kernel_walks_context', legacy), ('context_hoisted_gemm', hoisted)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name + '_us'] = round(e0.elapsed_time(e1) * 1e3 / 20, 1)
legacy()
a = out.clone()
b = hoisted()
torch.cuda.synchronize()
res['max_abs_diff'] = float((a - b).abs().max())
res['rows'] = B * J
res['flops_context_part'] = 2 * B * ctx.shape[1] * nl * 2 * 64
print(json.dumps(res))
