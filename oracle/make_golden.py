"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* from the UNMODIFIED reference forward.

Run in the build container (needs /root/reference):   python -m oracle.make_golden [preset ...]

Per preset (oracle/ref_loader.PRESETS) it writes
  tests/golden/calib_<preset>.npz   BatchNorm running stats of the calibrated synthetic checkpoint
  tests/golden/golden_<preset>.npz  reference outputs on the seeded synthetic batch (B=4, seed 1)
  tests/golden/spec_<preset>.json   state-dict names + shapes (the checkpoint contract)
The GPU box has no reference tree; tests rebuild the identical checkpoint from the seed + calib
fixture (oracle/synth_ckpt.py) and compare the CUDA path / the oracle against golden_*.npz.
"""
import json
import os
import sys

import numpy as np
import torch

from . import ref_loader as R
from . import synth_ckpt as S

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
CKPT_SEED = 0
CALIB_B, CALIB_SEED = 8, 123
TEST_B, TEST_SEED = 4, 1
FLOW_ROWS = 48


def make(preset):
    torch.set_num_threads(os.cpu_count())
    model, kw = R.build_reference(preset, seed=CKPT_SEED)
    template = model.state_dict()
    spec = {k: list(v.shape) for k, v in template.items()}

    model.load_state_dict(S.synth_state_dict(template, CKPT_SEED, None), strict=True)
    calib = S.calibrate_bn(model, R.synthetic_batch(CALIB_B, CALIB_SEED))
    sd = S.synth_state_dict(template, CKPT_SEED, calib)
    model.load_state_dict(sd, strict=True)
    model.eval()

    batch = R.synthetic_batch(TEST_B, TEST_SEED)
    with torch.no_grad():
        feats = model.backbone(batch['img'])
        out = model(batch)
    gold = {}
    for k in ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose', 'pred_pose6d', 'pred_pose_6d',
              'uncert_feat', 'body_feat2'):
        if k in out:
            gold[k] = out[k].numpy()
    if 'pred_segm_mask' in out:
        m = out['pred_segm_mask']
        st = 4 if m.shape[-1] >= 56 else 1
        gold['pred_segm_mask_sub'] = m[:, :, ::st, ::st].numpy()
        gold['segm_stride'] = np.array(st)
    assert out['log_phi'] is None and out['gt_pose_cond_idx'] == []
    # backbone features (sub-sampled) so a backbone bug is separable from a head bug
    if feats.dim() == 4:
        st = 8 if feats.shape[-1] >= 56 else 2
        gold['feat_sub'] = feats[:2, :, ::st, ::st].numpy()
        gold['feat_stride'] = np.array(st)
        gold['feat_absmean'] = feats.abs().mean(dim=(0, 2, 3)).numpy()
    else:
        gold['feat_sub'] = feats.numpy()
        gold['feat_stride'] = np.array(0)
    # non-degeneracy evidence: crop-to-crop differences must be >> tolerance
    gold['crop_diff_pose'] = np.array(float((out['pred_pose'][0] - out['pred_pose'][1]).abs().max()))
    gold['crop_diff_var'] = np.array(float((out['var_pose'][0] - out['var_pose'][1]).abs().max()))

    # RealNVP (runs only in training in the reference, nf_head.py:85-122; separately callable here)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(FLOW_ROWS, 9, generator=g) * 3.0
    z = torch.randn(FLOW_ROWS, 9, generator=g)
    with torch.no_grad():
        ctx_b = model.flow_head.cond_layer(out['uncert_feat'])          # [B,512]
        ctx = torch.repeat_interleave(ctx_b, FLOW_ROWS // TEST_B, dim=0)
        gold['flow_ctx'] = ctx_b.numpy()
        gold['flow_x'] = x.numpy()
        gold['flow_z'] = z.numpy()
        gold['flow_log_prob'] = model.flow_head.flow.log_prob(x, ctx).numpy()
        zb, ld = model.flow_head.flow.backward_p(x, ctx)
        gold['flow_backward_z'] = zb.numpy()
        gold['flow_logdet'] = ld.numpy()
        gold['flow_forward_x'] = model.flow_head.flow.forward_p(z, ctx).numpy()

    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, f'calib_{preset}.npz'), **calib)
    np.savez_compressed(os.path.join(GOLD, f'golden_{preset}.npz'), **gold)
    meta = {'kwargs': kw, 'ckpt_seed': CKPT_SEED, 'test_b': TEST_B, 'test_seed': TEST_SEED,
            'torch': torch.__version__, 'spec': spec}
    with open(os.path.join(GOLD, f'spec_{preset}.json'), 'w') as f:
        json.dump(meta, f)
    print(preset, 'feat|mean|', float(feats.abs().mean()), 'crop_diff_pose', float(gold['crop_diff_pose']),
          'crop_diff_var', float(gold['crop_diff_var']),
          'var range', float(out['var_pose'].min()), float(out['var_pose'].max()))


if __name__ == '__main__':
    for p in (sys.argv[1:] or list(R.PRESETS)):
        make(p)
