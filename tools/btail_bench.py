"""fused Bottleneck tail (poco_bottleneck_tail) against the two poco_conv launches it replaces, batch 256, CUDA events"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from poco_b200 import _lib as L
from poco_b200 import engine

B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 56
dev = 'cuda'
s = torch.cuda.current_stream().cuda_stream
a = engine.alloc_act(64, B, H, H, dev)
engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
r = engine.alloc_act(256, B, H, H, dev)
engine.act_view(r)[:, :, 1:H + 1, 1:H + 1].normal_()
m, o = engine.alloc_act(64, B, H, H, dev), engine.alloc_act(256, B, H, H, dev)
w2 = (torch.randn(9, 8, 64, 8, device=dev) * 0.05).half()
w3 = (torch.randn(1, 8, 256, 8, device=dev) * 0.1).half()
b2, b3 = torch.randn(64, device=dev) * 0.1, torch.randn(256, device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


for ctas in (0, 74):
    fused = L.BottleneckTail(a.desc(), o.desc(), r.ptr, r.plane_stride, w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), ctas, 0)
    c2 = L.Conv(a.desc(), m.desc(), w2.data_ptr(), b2.data_ptr(), None, 0, 3, 3, 1, 1, 1, 0, ctas, 0, None)
    c3 = L.Conv(m.desc(), o.desc(), w3.data_ptr(), b3.data_ptr(), r.ptr, r.plane_stride, 1, 1, 1, 0, 1, 0, ctas, 0, None)

    def two():
        L.run_op(c2, s)
        L.run_op(c3, s)
    t2 = timed(two)
    tf = timed(lambda: L.run_op(fused, s))
    gb = 2.0 * B * H * H * (64 + 256 + 256) / 1e9
    print(f'batch {B} max_ctas {ctas}: two launches {t2:.1f} us, fused {tf:.1f} us ({gb / tf * 1e6 / 1e3:.2f} TB/s of input + residual + output)', flush=True)
