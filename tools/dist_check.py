"""torchrun --nproc-per-node N tools/dist_check.py : sharded forward + ONE all-gather == unsharded forward (bitwise)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from common import build_model, synthetic_batch  # noqa: E402
from poco_b200 import dist as pdist  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
per = 4
model = build_model('cliff_w32', f'cuda:{local}')
full = synthetic_batch('cliff_w32', f'cuda:{local}', B=per * world)
lo, hi = pdist.shard_range(per * world, rank, world)
with torch.no_grad():
    mine = model.hot_path({k: v[lo:hi].contiguous() for k, v in full.items()})
    gathered = pdist.all_gather_outputs(mine)
    ok = True
    if rank == 0:
        ref = model.hot_path(full)
        for k in gathered:
            same = torch.equal(gathered[k], ref[k].reshape(gathered[k].shape))
            print(f'[dist_check] world={world} {k}: bitwise equal = {same}', flush=True)
            ok &= same
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
