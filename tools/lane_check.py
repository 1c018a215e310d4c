"""timing probe: eager vs CUDA-graph replay of the plan, with and without concurrent lanes; plus a direct
two-stream concurrency probe of capped conv launches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

from common import build_model, synthetic_batch  # noqa: E402
from poco_b200 import _lib as L  # noqa: E402
from poco_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for lanes in ('1', '0'):
    os.environ['POCO_B200_LANES'] = lanes
    m = build_model('cliff_w32', 'cuda')
    batch = synthetic_batch('cliff_w32', 'cuda', B=B)
    with torch.no_grad():
        m.use_cuda_graph = False
        m.hot_path(batch)
        eng = m._engine(B, batch['img'].device)
        t_eager = timeit(lambda: eng.plan.run())
        t_graph = timeit(lambda: eng.run(True))
    print(f'lanes={lanes} ops={eng.plan.num_ops} eager {t_eager:.3f} ms  graph {t_graph:.3f} ms', flush=True)
    del m, eng

# direct probe: two capped convs on two streams vs back to back on one
def mk(cin, H, cap):
    a = engine.alloc_act(cin, B, H, H, 'cuda')
    engine.act_view(a)[:, :, 1:H + 1, 1:H + 1].normal_()
    o = engine.alloc_act(cin, B, H, H, 'cuda')
    w = (torch.randn(9, cin // 8, cin, 8, device='cuda') * 0.05).half()
    b = torch.zeros(cin, device='cuda')
    d = L.Conv(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), None, 0, 3, 3, 1, 1, 1, 0, cap)
    return L.make_op(d), (a, o, w, b)


opA, keepA = mk(32, 56, 74)
opB, keepB = mk(256, 7, 74)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def seq():
    for _ in range(8):
        L.run_op(opA, s1.cuda_stream)
    for _ in range(8):
        L.run_op(opB, s1.cuda_stream)


def par():
    for _ in range(8):
        L.run_op(opA, s1.cuda_stream)
        L.run_op(opB, s2.cuda_stream)


for name, fn in (('sequential one stream', seq), ('two streams', par)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print(f'probe {name}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per 8+8 convs (cap 74 CTAs each)', flush=True)
