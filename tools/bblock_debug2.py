"""bring-up aid: is the two-launch block bit-stable across kernel configurations, and the fused one across runs?"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from gpu_util import run_basic_block, run_conv

tag = sys.argv[1]
H, W, N = 56, 56, 2
g = torch.Generator().manual_seed(1)
x = torch.randn(N, 32, H, W, generator=g)
w1, w2 = (torch.randn(32, 32, 3, 3, generator=g) * 0.08 for _ in range(2))
b1, b2 = (torch.randn(32, generator=g) * 0.2 for _ in range(2))
got = run_basic_block(x, w1, b1, w2, b2)
got2 = run_basic_block(x, w1, b1, w2, b2, 3)
mid = run_conv(x, w1, b1, relu=1)
two = run_conv(mid, w2, b2, relu=1, residual=x)
print(tag, 'fused run-to-run / grid-size mismatches', int((got != got2).sum()), 'fused vs two', int((got != two).sum()))
torch.save({'got': got, 'two': two, 'mid': mid}, f'/tmp/bb_{tag}.pt')
if tag != 'base':
    b = torch.load('/tmp/bb_base.pt')
    print(tag, 'two vs base two', int((two != b['two']).sum()), 'mid vs base mid', int((mid != b['mid']).sum()), 'got vs base got', int((got != b['got']).sum()))
