"""resident branch (poco_branch: the four BasicBlocks of the 128-channel 14x14 HRNet branch in one launch) against the
same blocks as eight poco_conv launches, batch 256, CUDA events"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from poco_b200 import _lib as L
from poco_b200 import engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 4
CH, H = 128, 14
dev = 'cuda'
s = torch.cuda.current_stream().cuda_stream
a = engine.alloc_act(CH, B, H, H, dev)
engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
m, o = engine.alloc_act(CH, B, H, H, dev), engine.alloc_act(CH, B, H, H, dev)
w = [(torch.randn(9, CH // 8, CH, 8, device=dev) * 0.02).half() for _ in range(2 * NB)]
b = [torch.randn(CH, device=dev) * 0.1 for _ in range(2 * NB)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
prof = torch.zeros(16, dtype=torch.int64, device=dev)
if os.environ.get('BR_PROF') == '1':
    os.environ['POCO_BRANCH_PROF'] = str(prof.data_ptr())


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


for ctas in (0, 74, 37, 20):
    br = L.Branch()
    br.in_, br.out, br.n_blocks, br.max_ctas = a.desc(), o.desc(), NB, ctas
    for i in range(2 * NB):
        br.weight[i], br.bias[i] = w[i].data_ptr(), b[i].data_ptr()
    convs = []
    src = a
    for k in range(NB):
        dst = o if k % 2 == 0 else a            # (timing only: the blocks ping-pong between two buffers)
        convs.append(L.Conv(src.desc(), m.desc(), w[2 * k].data_ptr(), b[2 * k].data_ptr(), None, 0, 3, 3, 1, 1, 1, 0, ctas, 0, None))
        convs.append(L.Conv(m.desc(), dst.desc(), w[2 * k + 1].data_ptr(), b[2 * k + 1].data_ptr(), src.ptr, src.plane_stride, 3, 3, 1, 1, 1, 0,
                            ctas, 0, None))
        src = dst

    def separate():
        for c in convs:
            L.run_op(c, s)
    t2 = timed(separate)
    tf = timed(lambda: L.run_op(br, s))
    fl = 2 * NB * 2.0 * B * H * H * CH * CH * 9
    print(f'{CH} ch {H}x{H} batch {B} x{NB} blocks max_ctas {ctas}: {2 * NB} conv launches {t2:.1f} us, resident branch {tf:.1f} us '
          f'({fl / tf / 1e6:.0f} TFLOP/s)', flush=True)
    if os.environ.get('BR_PROF') == '1':
        prof.zero_()
        L.run_op(br, s)
        torch.cuda.synchronize()
        v = prof.tolist()
        n = max(1, v[5])
        print('   issuer 0, cycles per crop: total %.0f | wait x_full %.0f act_ready %.0f w_full %.0f | issue %.0f' %
              (v[0] / n, v[1] / n, v[2] / n, v[3] / n, v[4] / n), flush=True)
        k = max(1, v[13])
        print('   epilogue warp 3, cycles per conv: total %.0f | wait acc_full %.0f | work %.0f' % (v[8] / k, v[9] / k, v[10] / k), flush=True)
