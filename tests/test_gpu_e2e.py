"""-m gpu: POCO.forward through the C ABI on a B200 against the reference-generated goldens."""
import numpy as np
import pytest
import torch

import emu
from common import GATED, PRESETS, build_model, load_preset, rel_err, synthetic_batch
from gpu_util import sync_or_die

pytestmark = pytest.mark.gpu

# The north-star gate (<= 1e-3 on the four gated outputs against the fp32 reference) is asserted on the parity mode,
# precision='split', in tests/test_gpu_split.py::test_parity_mode_meets_the_north_star_gate and in the
# benchmark-batch-size test below.  This file exercises the default precision='fp16' (BASELINE configs[1] "fp16"):
# fp16 weights + fp16 activation storage through ~110 sequential conv layers do NOT meet 1e-3 on pred_pose /
# pred_shape / pred_cam (measured 0.9-2.0e-2 / 2.0-3.5e-3 / 0.5-8.9e-3 on the calibrated synthetic checkpoint,
# DESIGN.md 4); the bounds below are what fp16 operands are expected to deliver, not the parity gate.
E2E_TOL = {'pred_pose': 6e-2, 'pred_shape': 6e-3, 'pred_cam': 1.5e-2, 'var_pose': 1e-3}
PARITY_TOL = 1e-3


@pytest.mark.parametrize('preset', PRESETS)
def test_forward_matches_reference_goldens(preset):
    meta, gold, _ = load_preset(preset)
    m = build_model(preset, 'cuda')
    batch = synthetic_batch(preset, 'cuda')
    with torch.no_grad():
        out = m(batch)
    sync_or_die(120)
    B = meta['test_b']
    # dict contract
    assert out['log_phi'] is None and out['gt_pose_cond_idx'] == []
    assert out['pred_pose'].shape == (B, 24, 3, 3) and out['var_pose'].shape == (B, 24)
    assert out['smpl_vertices'].shape == (B, 6890, 3) and out['pred_cam_t'].shape == (B, 3)
    if 'cliff' in preset:
        assert set(out) >= {'pred_pose_6d', 'uncert_feat', 'body_feat2', 'pred_fullimg_cam_t'}
    else:
        assert set(out) >= {'pred_pose6d', 'uncert_feat', 'pred_segm_mask'}
        assert out['pred_segm_mask'].shape[:2] == (B, 25)
    for k, v in out.items():
        if torch.is_tensor(v):
            assert v.dtype == torch.float32 and v.is_cuda, k
            assert torch.isfinite(v).all(), k
    errs = {k: rel_err(out[k].cpu().numpy(), gold[k]) for k in GATED}
    print(preset, errs)
    for k in GATED:
        assert errs[k] < E2E_TOL[k], (k, errs)
    # non-gated outputs are checked too (looser: features carry the raw fp16 noise)
    for k in ('uncert_feat', 'body_feat2', 'pred_pose6d', 'pred_pose_6d'):
        if k in gold:
            assert rel_err(out[k].cpu().numpy(), gold[k]) < 3e-2, k
    if 'pred_segm_mask_sub' in gold:
        st = int(gold['segm_stride'])
        assert rel_err(out['pred_segm_mask'][:, :, ::st, ::st].cpu().numpy(), gold['pred_segm_mask_sub']) < 5e-2


@pytest.mark.parametrize('preset,B,precision', [('cliff_w32', 256, 'fp16'), ('pare_w32', 128, 'fp16'),
                                                ('cliff_w32', 256, 'split'), ('pare_w32', 128, 'split')])
def test_benchmark_batch_sizes_equal_small_batches_and_goldens(preset, B, precision):
    """BASELINE configs[1] (POCO-CLIFF / HRNet-W32, 256 crops) and configs[2] (POCO-PARE / HRNet-W32, 128 crops) with
    the default schedule of that batch size (plan lanes with SM shares, m_group, half-size CTAs, 148-CTA persistent
    grids, CUDA-graph replay): crops 0-3 are the golden crops and must match the reference outputs; EVERY crop must
    equal the same crop run in a batch of 4 -- bitwise for everything a crop computes on its own (the whole backbone,
    the PARE head), to fp32 re-association for what goes through the split-K linear layers (their split count depends
    on the number of rows)."""
    meta, gold, _ = load_preset(preset)
    m = build_model(preset, 'cuda', precision=precision)
    small = synthetic_batch(preset, 'cuda')                     # the 4 golden crops
    extra = synthetic_batch(preset, 'cuda', B=B - meta['test_b'])
    big = {k: torch.cat([small[k], extra[k] + (0.25 if k == 'img' else 0.0)], 0).contiguous() for k in small}
    with torch.no_grad():
        for _ in range(3):              # eager warm-up, graph capture, replay
            out = m.hot_path(big)
    sync_or_die(180)
    tol = E2E_TOL if precision == 'fp16' else {k: PARITY_TOL for k in GATED}
    errs = {k: rel_err(out[k][:meta['test_b']].cpu().numpy(), gold[k]) for k in GATED}
    print(preset, B, precision, errs)
    for k in GATED:
        assert torch.isfinite(out[k]).all(), k
        assert errs[k] < tol[k], (k, errs)
    exact = ('uncert_feat',) if 'cliff' in preset else ('uncert_feat', 'pred_pose', 'pred_shape', 'pred_cam', 'pred_segm_mask')
    with torch.no_grad():
        for i0 in range(0, B, 4):
            sub = m.hot_path({k: v[i0:i0 + 4].contiguous() for k, v in big.items()})
            for k in exact:
                assert torch.equal(sub[k], out[k][i0:i0 + 4]), (k, i0)
            for k in GATED:
                assert rel_err(sub[k].cpu().numpy(), out[k][i0:i0 + 4].cpu().numpy()) < 1e-5, (k, i0)
    sync_or_die(180)


@pytest.mark.parametrize('preset', ['pare_r50', 'cliff_w32'])
def test_gpu_equals_cpu_replay_of_the_same_schedule(preset):
    """the CUDA kernels and the CPU op interpreter implement the same arithmetic (fp16 storage, fp32
    accumulate); they differ by accumulation order only, but a flipped fp16 rounding early in the
    network propagates like any other fp16 rounding, so the bound is the fp16-vs-fp32 one"""
    meta, gold, _ = load_preset(preset)
    m = build_model(preset, 'cuda')
    with torch.no_grad():
        out = m.hot_path(synthetic_batch(preset, 'cuda'))
    sync_or_die(120)
    mc = build_model(preset)
    batch = synthetic_batch(preset)
    eng = mc._build_engine(meta['test_b'], torch.device('cpu'))
    eng.img.copy_(batch['img'])
    if eng.bbox is not None:
        eng.bbox.copy_(batch['bbox_info'])
    emu.run_plan_ops(eng.plan.ops, eng.plan.keep)
    for k in GATED:
        e = rel_err(out[k].cpu().numpy(), eng.out[k].reshape(out[k].shape).numpy())
        print(preset, k, e)
        assert e < E2E_TOL[k], (k, e)


def test_cuda_graph_replay_is_bitwise_equal_to_eager():
    batch = synthetic_batch('cliff_w32', 'cuda')
    m = build_model('cliff_w32', 'cuda', use_cuda_graph=False)
    g = build_model('cliff_w32', 'cuda', use_cuda_graph=True)
    with torch.no_grad():
        a = m.hot_path(batch)
        for _ in range(3):          # eager warm-up, capture, replay
            b = g.hot_path(batch)
    sync_or_die(120)
    for k in GATED:
        assert torch.equal(a[k], b[k]), k


def test_outputs_are_fresh_and_batch_invariant():
    m = build_model('cliff_w32', 'cuda')
    batch = synthetic_batch('cliff_w32', 'cuda')
    with torch.no_grad():
        a = m.hot_path(batch)
        b = m.hot_path(batch)
        assert a['pred_pose'].data_ptr() != b['pred_pose'].data_ptr()       # caller owns the outputs
        assert torch.equal(a['pred_pose'], b['pred_pose'])                 # deterministic
        # a crop's result does not depend on the batch it travels in (bitwise shard == single-GPU)
        one = m.hot_path({k: v[1:2] for k, v in batch.items()})
    sync_or_die(120)
    for k in GATED:
        assert torch.equal(one[k][0], a[k][1]), k


def test_debug_conv_kernels_agree_with_tcgen05_path():
    batch = synthetic_batch('pare_r50', 'cuda')
    import os
    os.environ['POCO_B200_CONV_IMPL'] = '1'
    try:
        ref = build_model('pare_r50', 'cuda')
    finally:
        os.environ.pop('POCO_B200_CONV_IMPL')
    tc = build_model('pare_r50', 'cuda')
    with torch.no_grad():
        a, b = ref.hot_path(batch), tc.hot_path(batch)
    sync_or_die(120)
    for k in GATED:
        assert rel_err(b[k].cpu().numpy(), a[k].cpu().numpy()) < E2E_TOL[k], k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_shards_plus_one_all_gather_equal_single_gpu():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(root, 'tools', 'dist_check.py')],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_ragged_batch_sizes_and_reload():
    """B = 1, an odd batch, a second batch size on the same model, and a weight reload (plans are rebuilt)"""
    meta, gold, sd = load_preset('cliff_w32')
    m = build_model('cliff_w32', 'cuda')
    full = synthetic_batch('cliff_w32', 'cuda', B=7)
    with torch.no_grad():
        o7 = m.hot_path(full)
        o1 = m.hot_path({k: v[:1] for k, v in full.items()})
        o3 = m.hot_path({k: v[2:5].contiguous() for k, v in full.items()})
    sync_or_die(120)
    for k in GATED:
        assert torch.equal(o1[k][0], o7[k][0]) and torch.equal(o3[k], o7[k][2:5]), k
    # reloading weights invalidates the prepared plans: doubling a bias must change the output
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2['head.deccam.bias'] = sd2['head.deccam.bias'] + 1.0
    m.load_state_dict(sd2)
    with torch.no_grad():
        o7b = m.hot_path(full)
    sync_or_die(120)
    assert (o7b['pred_cam'] - o7['pred_cam']).abs().min() > 0.5


def test_empty_batch_and_bad_inputs_raise():
    m = build_model('pare_r50', 'cuda')
    with pytest.raises(ValueError):
        m({'img': torch.zeros(2, 3, 128, 128, device='cuda')})
    with pytest.raises(KeyError):
        build_model('cliff_w32', 'cuda')({'img': torch.zeros(1, 3, 224, 224, device='cuda')})     # bbox_info missing
    m.train()
    with pytest.raises(Exception):
        m({'img': torch.zeros(1, 3, 224, 224, device='cuda')})


def test_stream_runner_equals_the_separate_steps():
    """f2: StreamRunner.step == oracle crop -> POCO.forward -> oracle uncertainty post-processing"""
    from common import build_model
    from oracle import crop_oracle as C
    from oracle import uncert_oracle as U
    from poco_b200 import StreamRunner
    m = build_model('cliff_w32', 'cuda')
    frame = C.synthetic_frame(2, 720, 1280)
    boxes = C.synthetic_boxes(2, 5, 720, 1280)
    run = StreamRunner(m, bbox_scale=1.1)
    out = run.step(torch.from_numpy(frame).cuda(), torch.from_numpy(boxes.astype(np.float32)))
    torch.cuda.synchronize()
    ref_batch = {k: torch.from_numpy(v).cuda() for k, v in C.crop_batch(frame, boxes, 1.1).items()}
    with torch.no_grad():
        ref = m(ref_batch)
    for k in ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose'):
        assert torch.equal(out[k], ref[k]), k                   # identical crops -> identical forward
    p = U.prepare_uncert(ref['var_pose'].cpu().numpy())
    _, gl = U.global_uncert(p, 'hrnet_w32-cliff')
    assert np.array_equal(out['variance'].cpu().numpy(), p)
    assert np.array_equal(out['variance_global'].cpu().numpy(), np.clip(gl, 0, 0.99))      # tester.py:245
    raw = StreamRunner(m, bbox_scale=1.1, clip_global=False).step(torch.from_numpy(frame).cuda(),
                                                                  torch.from_numpy(boxes.astype(np.float32)))
    assert np.array_equal(raw['variance_global'].cpu().numpy(), gl)                        # tester.py:418-421
    assert out['orig_cam'].shape == (5, 4) and torch.isfinite(out['orig_cam']).all()


def test_latency_mode_chains_match_the_default_schedule():
    """POCO(latency_mode=True): small batches run every HRNet branch as one persistent chained launch; same results
    as the conv-by-conv schedule up to the fp32 accumulation order (one fp16 rounding per layer can flip)"""
    from common import build_model
    from poco_b200 import _lib as L
    a = build_model('cliff_w32', 'cuda')
    b = build_model('cliff_w32', 'cuda', latency_mode=True)
    batch = synthetic_batch('cliff_w32', 'cuda')
    with torch.no_grad():
        oa, ob = a.hot_path(batch), b.hot_path(batch)
        ob2 = b.hot_path(batch)
    sync_or_die(120)
    kinds = [op.kind for op in b._engine(4, batch['img'].device).plan.ops]
    assert L.OP_CONV_CHAIN in kinds and L.OP_CONV_CHAIN not in [op.kind for op in a._engine(4, batch['img'].device).plan.ops]
    meta, gold, _ = load_preset('cliff_w32')
    for k in GATED:
        assert torch.equal(ob[k], ob2[k]), k
        assert rel_err(ob[k].cpu().numpy(), oa[k].cpu().numpy()) < E2E_TOL[k], k
        assert rel_err(ob[k].cpu().numpy(), gold[k]) < E2E_TOL[k], k


def test_stream_runner_graph_replay_and_padded_buckets():
    """f2: the per-frame CUDA graph (crop -> plan -> uncertainty -> cameras) replays bit-identically to the eager body,
    and a detection count between buckets (11 -> 16, padded with copies of the first box) returns the 11 real rows"""
    from common import build_model
    from oracle import crop_oracle as C
    from poco_b200 import StreamRunner
    m = build_model('cliff_w32', 'cuda')
    eager, graphed = StreamRunner(m, graph=False), StreamRunner(m, graph=True)
    for seed, n in ((3, 4), (4, 11), (5, 4), (6, 11)):          # second visit of a bucket = pure replay with new inputs
        frame = torch.from_numpy(C.synthetic_frame(seed, 720, 1280)).cuda()
        boxes = torch.from_numpy(C.synthetic_boxes(seed, n, 720, 1280).astype(np.float32))
        a = eager.step(frame, boxes)
        b = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in graphed.step(frame, boxes).items()}
        torch.cuda.synchronize()
        assert b['pred_pose'].shape[0] == n and b['confidence'].shape == (n,)
        for k in ('uncert_feat', 'variance', 'variance_global', 'orig_cam') + GATED:
            if n == m.bucket(n):
                assert torch.equal(a[k], b[k]), (k, n)
            else:       # another plan batch size: the split-K linear layers re-associate (crops themselves are exact)
                assert rel_err(b[k].cpu().numpy(), a[k].cpu().numpy()) < 1e-5, (k, n)
        assert torch.equal(a['uncert_feat'], b['uncert_feat'])
    assert len(graphed._graphs) == 2
