"""Pins the oracle (oracle/poco_oracle.py) to the golden vectors generated from the UNMODIFIED
reference forward (oracle/make_golden.py).  fp32 vs fp32: only summation order may differ."""
import numpy as np
import pytest
import torch

from common import PRESETS, load_preset, rel_err, synthetic_batch
from oracle import poco_oracle as O

TOL = 2e-5


@pytest.mark.parametrize('preset', PRESETS)
def test_forward_matches_reference(preset):
    meta, gold, sd = load_preset(preset)
    bb, head = meta['kwargs']['backbone'].split('-')
    batch = synthetic_batch(preset)
    with torch.no_grad():
        out = O.poco_forward(batch, sd, bb, head, meta['kwargs']['uncert_inp_type'])
    assert out['log_phi'] is None and out['gt_pose_cond_idx'] == []
    for k in ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose', 'pred_pose6d', 'pred_pose_6d', 'uncert_feat', 'body_feat2'):
        if k in gold:
            assert out[k].shape == gold[k].shape, k
            assert rel_err(out[k].numpy(), gold[k]) < TOL, k
    if 'pred_segm_mask_sub' in gold:
        st = int(gold['segm_stride'])
        assert rel_err(out['pred_segm_mask'][:, :, ::st, ::st].numpy(), gold['pred_segm_mask_sub']) < TOL


@pytest.mark.parametrize('preset', ['pare_w32', 'cliff_w48cls'])
def test_backbone_features_match_reference(preset):
    meta, gold, sd = load_preset(preset)
    bb = meta['kwargs']['backbone'].split('-')[0]
    with torch.no_grad():
        f = O.backbone(synthetic_batch(preset)['img'], sd, bb)
    st = int(gold['feat_stride'])
    sub = f[:2, :, ::st, ::st] if st else f
    assert rel_err(sub.numpy(), gold['feat_sub']) < TOL


@pytest.mark.parametrize('preset', ['pare_w32', 'cliff_w32'])
def test_realnvp_matches_reference(preset):
    meta, gold, sd = load_preset(preset)
    rows = gold['flow_x'].shape[0]
    ctx = torch.repeat_interleave(torch.from_numpy(gold['flow_ctx']), rows // meta['test_b'], 0)
    x, z = torch.from_numpy(gold['flow_x']), torch.from_numpy(gold['flow_z'])
    with torch.no_grad():
        zb, ld = O.realnvp_backward(x, ctx, sd)
        assert rel_err(zb.numpy(), gold['flow_backward_z']) < TOL
        assert rel_err(ld.numpy(), gold['flow_logdet']) < TOL
        assert rel_err(O.realnvp_log_prob(x, ctx, sd).numpy(), gold['flow_log_prob']) < TOL
        fx = O.realnvp_forward(z, ctx, sd)
        assert rel_err(fx.numpy(), gold['flow_forward_x']) < TOL
        # property: forward_p(backward_p(x)) == x  (round-trips to 3e-7)
        assert rel_err(O.realnvp_forward(zb, ctx, sd).numpy(), gold['flow_x']) < 1e-5


def test_rot6d_is_a_rotation():
    x = torch.randn(64, 6, generator=torch.Generator().manual_seed(3))
    R = O.rot6d_to_rotmat(x)
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R.transpose(1, 2) @ R, eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)


def test_golden_is_not_degenerate():
    """the calibrated synthetic checkpoint must make outputs differ between crops by >> tolerance"""
    for p in PRESETS:
        _, gold, _ = load_preset(p)
        assert float(gold['crop_diff_pose']) > 0.1 and float(gold['crop_diff_var']) > 5e-3


def test_crop_oracle_matches_reference_golden():
    """the numpy restatement of get_single_image_crop_demo / calculate_bbox_info equals the
    outputs of the reference functions (tests/golden/crop_golden.npz, oracle/make_golden_crop.py) bit for bit"""
    import os

    import numpy as np

    from oracle import crop_oracle as C
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crop_golden.npz'))
    assert np.array_equal(g['frame'], C.synthetic_frame(0))
    assert np.array_equal(g['boxes'], C.synthetic_boxes(0).astype(np.float32))
    o = C.crop_batch(g['frame'], g['boxes'].astype(np.float64), float(g['scale']))
    assert np.array_equal(o['img'], g['img'])
    assert np.array_equal(o['bbox_info'], g['bbox_info'])
    assert np.array_equal(o['focal_length'], g['focal_length'])
    # sanity of the fixture: crops differ from each other and include zero-filled borders
    assert np.abs(g['img'][0] - g['img'][1]).max() > 1.0
    assert (g['img'][1][0] == g['img'][1][0, 0, 0]).mean() > 0.05       # box 1 hangs over the frame corner


def test_uncert_oracle_matches_reference_golden():
    """prepare_uncert / get_global_uncert restatement == the reference functions' outputs"""
    import os

    import numpy as np

    from oracle import uncert_oracle as U
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'uncert_golden.npz'))
    for bb in ('cliff', 'pare'):
        for kin in (0, 1):
            tag = f'{bb}_{kin}'
            p = U.prepare_uncert(g['var'], bool(kin))
            t, gl = U.global_uncert(p, bb)
            assert np.array_equal(p, g['prepared_' + tag]) and np.array_equal(t, g['thresholded_' + tag])
            assert np.array_equal(gl, g['global_' + tag])
    assert (g['thresholded_cliff_0'] == 1.0).all(1).sum() < (g['thresholded_pare_0'] == 1.0).all(1).sum()
