"""Does running the batch as several independent sub-batches on concurrent streams hide the fixed per-launch cost
(prologue + drain, ~10 us per conv: profiles/r02b_conv_slope_fixed_vs_marginal.csv)?
usage: python tools/split_batch_bench.py <tag> <ways> [B=256] [preset]   (environment knobs apply)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

tag, ways = sys.argv[1], int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
preset = sys.argv[4] if len(sys.argv) > 4 else 'cliff_w32'
dev = torch.device('cuda', 0)
model, sd, meta = bench.load_model_and_sd(preset)
model = model.to(dev).eval()
sub = B // ways
with torch.no_grad():
    engs = [model._build_engine(sub, dev) for _ in range(ways)]
    batch = bench.build_inputs(preset, B, dev)
    for i, e in enumerate(engs):
        e.img.copy_(batch['img'][i * sub:(i + 1) * sub])
        if e.bbox is not None:
            e.bbox.copy_(batch['bbox_info'][i * sub:(i + 1) * sub])
    streams = [torch.cuda.Stream(device=dev) for _ in range(ways)]

    def run_all():
        cur = torch.cuda.current_stream()
        for s, e in zip(streams, engs):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                e.plan.run(s.cuda_stream)
        for s in streams:
            cur.wait_stream(s)

    run_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run_all()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    e0.record()
    for _ in range(K):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f'{tag},{preset},{B},{ways},{ms:.3f},{B / ms * 1e3:.0f}', flush=True)
