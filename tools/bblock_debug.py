"""bring-up aid: where does the fused BasicBlock differ from the two-launch block?"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from gpu_util import conv_reference, run_basic_block, run_conv

H, W, N = 56, 56, 2
g = torch.Generator().manual_seed(1)
x = torch.randn(N, 32, H, W, generator=g)
w1, w2 = (torch.randn(32, 32, 3, 3, generator=g) * 0.08 for _ in range(2))
b1, b2 = (torch.randn(32, generator=g) * 0.2 for _ in range(2))
for name, (wa, wb, ba, bb) in {'full': (w1, w2, b1, b2),
                               'conv2=identity': (w1, None, b1, None),
                               'conv1=identity': (None, w2, None, b2)}.items():
    ident = torch.zeros(32, 32, 3, 3)
    for i in range(32):
        ident[i, i, 1, 1] = 1.0
    wa = ident if wa is None else wa
    wb = ident if wb is None else wb
    ba = torch.zeros(32) if ba is None else ba
    bb = torch.zeros(32) if bb is None else bb
    got = run_basic_block(x, wa, ba, wb, bb)
    mid = run_conv(x, wa, ba, relu=1)
    two = run_conv(mid, wb, bb, relu=1, residual=x)
    ref = conv_reference(conv_reference(x, wa, ba, relu=1).half().float(), wb, bb, relu=1, residual=x)
    d = (got - two).abs()
    bad = (d > 0).nonzero()
    print(name, 'mismatch', int((d > 0).sum()), 'of', d.numel(), 'max', float(d.max()),
          'err fused vs ref', float((got - ref).abs().max()), 'two vs ref', float((two - ref).abs().max()))
    if len(bad):
        print('  first', bad[:8].tolist())
        q = (bad[:, 0] * 58 * 58 + (bad[:, 2] + 1) * 58 + bad[:, 3] + 1)
        print('  q mod 128 histogram (top):', torch.bincount(q % 128, minlength=128).topk(5))
        print('  q mod 384 min/max', int((q % 384).min()), int((q % 384).max()))
