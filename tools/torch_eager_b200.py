"""Stock-PyTorch-eager comparison on the same B200 ('second comparison'): the reference's forward is
plain torch ops (cuDNN convs, cuBLAS linears, separate BN / ReLU / add / upsample kernels); the reference tree
cannot travel to the GPU box, so its pinned functional restatement (oracle/poco_oracle.py -- the same torch calls
in the same order, checked against the reference forward) is moved to the device and timed with CUDA events.
MEASUREMENT AID ONLY: reported next to bench.py's numbers, never part of the product path.
    python tools/torch_eager_b200.py [preset] [batch] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import poco_oracle as O  # noqa: E402
from synth import ckpt as S  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else 'cliff_w32'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
gd = os.path.join(ROOT, 'tests', 'golden')
meta = json.load(open(os.path.join(gd, f'spec_{preset}.json')))
sd = S.synth_state_dict(S.template_from_spec(meta, 0), 0, np.load(os.path.join(gd, f'calib_{preset}.npz')))
sd = {k: v.cuda() for k, v in sd.items()}
bb, head = meta['kwargs']['backbone'].split('-')
uit = meta['kwargs']['uncert_inp_type']
batch = S.synthetic_batch(B, 1, 'cuda')
res = {'preset': preset, 'batch': B, 'steps': steps, 'torch': torch.__version__, 'cudnn': torch.backends.cudnn.version()}


def timed(tag, ctx):
    with torch.no_grad(), ctx:
        for _ in range(3):
            out = O.poco_forward(batch, sd, bb, head, uit)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = O.poco_forward(batch, sd, bb, head, uit)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res[tag] = {'ms_per_step': round(ms, 3), 'crops_per_s': round(B / ms * 1e3, 1)}
    return out


import contextlib  # noqa: E402

res['flags'] = {'cudnn.benchmark': torch.backends.cudnn.benchmark, 'cudnn.allow_tf32': torch.backends.cudnn.allow_tf32,
                'matmul.allow_tf32': torch.backends.cuda.matmul.allow_tf32}
timed('fp32_torch_defaults', contextlib.nullcontext())
timed('autocast_fp16', torch.autocast('cuda', dtype=torch.float16))
batch_cl = dict(batch, img=batch['img'].contiguous(memory_format=torch.channels_last))
batch = batch_cl
timed('autocast_fp16_channels_last_input', torch.autocast('cuda', dtype=torch.float16))
print(json.dumps(res))
