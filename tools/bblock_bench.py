"""fused BasicBlock (poco_basic_block) against the same block as two poco_conv launches, batch 256, CUDA events"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from poco_b200 import _lib as L
from poco_b200 import engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H = int(sys.argv[3]) if len(sys.argv) > 3 else (56 if CH == 32 else 28)
dev = 'cuda'
s = torch.cuda.current_stream().cuda_stream
a = engine.alloc_act(CH, B, H, H, dev)
engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
m, o = engine.alloc_act(CH, B, H, H, dev), engine.alloc_act(CH, B, H, H, dev)
w = [(torch.randn(9, CH // 8, CH, 8, device=dev) * 0.05).half() for _ in range(2)]
b = [torch.randn(CH, device=dev) * 0.1 for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
prof = torch.zeros(16, dtype=torch.int64, device=dev)
if os.environ.get('BB_PROF') == '1':
    os.environ['POCO_BBLOCK_PROF'] = str(prof.data_ptr())


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


for ctas in (0, 148, 74, 37):
    fused = L.BasicBlock(a.desc(), o.desc(), w[0].data_ptr(), b[0].data_ptr(), w[1].data_ptr(), b[1].data_ptr(), ctas, 0)
    c1 = L.Conv(a.desc(), m.desc(), w[0].data_ptr(), b[0].data_ptr(), None, 0, 3, 3, 1, 1, 1, 0, ctas, 0, None)
    c2 = L.Conv(m.desc(), o.desc(), w[1].data_ptr(), b[1].data_ptr(), a.ptr, a.plane_stride, 3, 3, 1, 1, 1, 0, ctas, 0, None)

    def two():
        L.run_op(c1, s)
        L.run_op(c2, s)
    t2 = timed(two)
    tf = timed(lambda: L.run_op(fused, s))
    print(f'{CH} ch {H}x{H} batch {B} max_ctas {ctas}: two launches {t2:.1f} us, fused {tf:.1f} us', flush=True)
    if os.environ.get('BB_PROF') == '1':
        prof.zero_()
        L.run_op(fused, s)
        torch.cuda.synchronize()
        if CH == 64 and not os.environ.get('POCO_B200_BLOCK64_RESIDENT'):       # bblock64_tc.cu: another counter layout
            v = prof.tolist()
            n = max(1, v[6])
            print('   conv1 issuer, cycles per unit: total %.0f | wait in_full %.0f acc1_free %.0f | issue %.0f   (units/cta %.1f)' %
                  (v[0] / n, v[1] / n, v[2] / n, v[4] / n, v[6] / max(1, v[7])), flush=True)
            print('   conv2 issuer, cycles per unit: total %.0f | wait mid_full %.0f acc2_free %.0f w2_full %.0f | issue %.0f' %
                  (v[8] / n, v[9] / n, v[10] / n, v[11] / n, v[12] / n), flush=True)
            continue
        for w_ in range(2):
            v = prof.tolist()[w_ * 8:w_ * 8 + 8]
            n = max(1, v[6])
            print('   conv%d issuer, cycles per unit: total %.0f | wait in_full %.0f acc1_free %.0f mid_full %.0f acc2_free %.0f | issue %.0f   (units/cta %.1f)' %
                  (w_ + 1, v[0] / n, v[1] / n, v[2] / n, v[3] / n, v[4] / n, v[5] / n, v[6] / max(1, v[7])), flush=True)
