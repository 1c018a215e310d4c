"""BASELINE configs[4]: 1080p video streams, one stream per GPU, D detections per frame through StreamRunner
(crop -> POCO-CLIFF/HRNet-W32 -> uncertainty post-processing -> original-image cameras; with --smpl also the device
mesh stage), frames resident on the device, one device -> host read of the per-frame confidence.

  python tools/stream_bench.py [D] [frames] [--smpl]                                  # one stream on cuda:0
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \\
      tools/stream_bench.py 8 200                                                     # 8 streams, one per GPU

Streams are independent (no collective on the data path); under torchrun the ranks meet at a barrier before and
after the timed region and the aggregate is frames of all streams / max-over-ranks time.  Prints one text line
(rank 0) and one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from common import build_model  # noqa: E402
from synth import frames as C  # noqa: E402
from poco_b200 import StreamRunner  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
with_smpl = '--smpl' in sys.argv[1:]
D = int(args[0]) if len(args) > 0 else 8
F = int(args[1]) if len(args) > 1 else 200
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
kw = {}
if with_smpl:
    from synth import smpl_model
    kw['smpl_model'] = smpl_model.synthetic_model(0)
kw['latency_mode'] = '--no-chains' not in sys.argv[1:]
m = build_model('cliff_w32', dev, **kw)
run = StreamRunner(m, graph='--no-graph' not in sys.argv[1:])
frames = [torch.from_numpy(C.synthetic_frame(s + 4 * rank, 1080, 1920)).to(dev) for s in range(4)]
boxes = torch.from_numpy(C.synthetic_boxes(1 + rank, D, 1080, 1920).astype(np.float32)).to(dev)
for i in range(10):
    run.step(frames[i % 4], boxes)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


barrier()
lat = []
t0 = time.perf_counter()
for i in range(F):
    t1 = time.perf_counter()
    out = run.step(frames[i % 4], boxes)
    conf = out['confidence'].cpu()          # the per-frame result a caller reads back (synchronises the frame)
    lat.append(time.perf_counter() - t1)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
lat.sort()
t = torch.tensor([dt], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
barrier()
dt = float(t.item())
if rank == 0:
    print(f'stream: {world} stream(s), {D} detections/frame, {F} frames each: {world * F / dt:.1f} fps aggregate '
          f'({F / dt:.1f} per stream, {world * F * D / dt:.0f} crops/s), {dt / F * 1e3:.2f} ms/frame')
    print(json.dumps({'metric': 'frames/sec', 'value': round(world * F / dt, 1), 'per_stream_fps': round(F / dt, 1),
                      'n_gpus': world, 'streams': world, 'detections_per_frame': D, 'frames_per_stream': F,
                      'ms_per_frame': round(dt / F * 1e3, 3), 'frame': '1080p uint8 RGB resident on the device',
                      'frame_latency_ms': {'p50': round(lat[len(lat) // 2] * 1e3, 3), 'p99': round(lat[int(len(lat) * 0.99)] * 1e3, 3),
                                           'note': 'rank 0: step() + read-back of the confidence, wall clock'},
                      'cuda_graph_per_frame': run.use_graph,
                      'smpl_mesh_stage': with_smpl, 'timing': 'wall clock between barriers, max over ranks'}))
if world > 1:
    dist.destroy_process_group()
