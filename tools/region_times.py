"""Where does the step go?  Replays the plan with its lanes on real streams and times every region
between fork / join points (CUDA events on the main stream).  usage: python tools/region_times.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import torch  # noqa: E402

from poco_b200 import _lib as L  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda', 0)
model, sd, meta = bench.load_model_and_sd('cliff_w32')
model = model.to(dev).eval()
batch = bench.build_inputs('cliff_w32', B, dev)
with torch.no_grad():
    model.hot_path(batch)
eng = model._engine(B, dev)
ops = eng.plan.ops
main = torch.cuda.current_stream()
sides = [None] + [torch.cuda.Stream() for _ in range(7)]


def replay(marks):
    for i, op in enumerate(ops):
        if op.kind == L.OP_FORK:
            e = torch.cuda.Event(enable_timing=True)
            e.record(main)
            marks.append((i, 'fork', e, None))
            for k in range(1, op.u.sync.n_lanes):
                sides[k].wait_event(e)
            continue
        if op.kind == L.OP_JOIN:
            lane_ends = []
            e0_ = torch.cuda.Event(enable_timing=True)
            e0_.record(main)
            lane_ends.append(e0_)
            for k in range(1, op.u.sync.n_lanes):
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(sides[k])
                main.wait_event(ev)
                lane_ends.append(ev)
            e = torch.cuda.Event(enable_timing=True)
            e.record(main)
            marks.append((i, 'join', e, lane_ends))
            continue
        s = main if op.lane == 0 else sides[op.lane]
        L.run_op(op, s.cuda_stream)


for _ in range(2):
    replay([])
torch.cuda.synchronize()
marks = []
e0 = torch.cuda.Event(enable_timing=True)
e0.record(main)
replay(marks)
e1 = torch.cuda.Event(enable_timing=True)
e1.record(main)
torch.cuda.synchronize()
print(f'total {e0.elapsed_time(e1):.3f} ms  ({len(ops)} ops)')
prev, prev_i, prev_kind = e0, 0, 'start'
for i, kind, e, lane_ends in marks + [(len(ops), 'end', e1, None)]:
    n = sum(1 for o in ops[prev_i:i] if o.kind not in (L.OP_FORK, L.OP_JOIN))
    what = 'lanes' if prev_kind == 'fork' else 'serial'
    labels = {}
    for o in ops[prev_i:i]:
        if o.kind in (L.OP_FORK, L.OP_JOIN):
            continue
        lab = bench.op_label(o)
        labels[lab] = labels.get(lab, 0) + 1
    top = ', '.join(f'{v}x {k}' for k, v in sorted(labels.items(), key=lambda kv: -kv[1])[:4])
    lanes_txt = ''
    if lane_ends:
        lanes_txt = '  lanes end at ' + ' '.join(f'{prev.elapsed_time(le):.3f}' for le in lane_ends)
    print(f'ops {prev_i:4d}-{i:4d} {what:6s} {n:3d} ops {prev.elapsed_time(e):8.3f} ms   {top}{lanes_txt}')
    prev, prev_i, prev_kind = e, i, kind
