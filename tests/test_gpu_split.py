"""-m gpu: the split-precision ("parity") mode.  Every activation / weight is a hi + lo fp16 pair and every conv runs
x_hi W_hi + x_lo W_hi + x_hi W_lo in one fp32 tcgen05 accumulation chain, so the CUDA path reproduces the reference's
fp32 forward (pocolib/core/config.py:154 PRECISION=32; pocolib/models/poco.py:99-129) to the north star's 1e-3 --
measured ~1e-5 -- on every gated output of every preset."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import GATED, PRESETS, build_model, load_preset, rel_err, synthetic_batch
from gpu_util import conv_reference_split, run_conv, split16, stream, sync_or_die
from oracle import poco_oracle as O
from poco_b200 import _lib as L
from poco_b200 import engine
from test_gpu_ops import _case_tensors

pytestmark = pytest.mark.gpu

# the north-star gate (BASELINE.json): max-abs-err / max-abs-ref <= 1e-3 on each gated output, fp32 reference
PARITY_TOL = 1e-3
# one conv against float64 arithmetic on the same hi + lo operands: dropped x_lo W_lo term (2^-22), fp32 accumulation
# in the tensor core, fp16 subnormal flush of tiny lo parts, hi + lo storage of the result (2^-21)
SPLIT_CONV_TOL = 2e-5

# (Cin, Cout, k, stride, H, N, residual, relu)
SPLIT_CASES = [
    (32, 32, 3, 1, 56, 2, True, 1), (64, 64, 3, 1, 28, 2, True, 1), (16, 32, 3, 1, 8, 1, False, 0),
    (48, 48, 3, 1, 56, 1, True, 1), (64, 256, 1, 1, 56, 1, True, 1), (256, 64, 1, 1, 56, 1, False, 1),
    (128, 128, 3, 1, 14, 3, True, 1), (256, 256, 3, 1, 7, 5, True, 1), (480, 256, 3, 1, 28, 1, False, 1),
    (256, 256, 3, 1, 56, 1, False, 1), (384, 384, 3, 1, 7, 2, True, 1), (1024, 2048, 1, 1, 7, 2, False, 1),
    (32, 256, 1, 1, 28, 2, False, 0),
    # gather mode
    (3, 64, 3, 2, 224, 1, False, 1), (64, 64, 3, 2, 112, 1, False, 1), (32, 64, 3, 2, 56, 2, False, 1),
    (256, 256, 3, 2, 14, 2, True, 2), (3, 64, 7, 2, 224, 1, False, 1), (256, 512, 1, 2, 56, 1, False, 0),
    (128, 256, 3, 2, 14, 3, True, 1), (32, 32, 3, 2, 56, 3, False, 1), (64, 128, 3, 2, 28, 5, True, 1),
]


@pytest.mark.parametrize('case', SPLIT_CASES, ids=lambda c: 'c%d-%d_k%d_s%d_h%d_n%d' % c[:6])
def test_conv_split_precision(case):
    cin, cout, k, stride, H, N, res, relu = case
    x, w, b, r = _case_tensors(cin, cout, k, stride, H, N, res)
    ref = conv_reference_split(x, w, b, stride, None, relu, r)
    out = run_conv(x, w, b, stride, None, relu, r, impl=0, split=True)
    e = rel_err(out.numpy(), ref.numpy())
    assert e < SPLIT_CONV_TOL, e
    # and it is a real improvement over one fp16 rounding of the operands (2^-11)
    assert e < 0.05 * rel_err(run_conv(x, w, b, stride, None, relu, r, impl=0).numpy(), ref.numpy())


def test_conv_split_debug_kernel_and_batch_invariance():
    x, w, b, r = _case_tensors(64, 64, 3, 1, 28, 4, True)
    ref = conv_reference_split(x, w, b, 1, None, 1, r)
    assert rel_err(run_conv(x, w, b, 1, None, 1, r, impl=1, split=True).numpy(), ref.numpy()) < SPLIT_CONV_TOL
    full = run_conv(x, w, b, 1, None, 1, r, split=True)
    for i in range(4):
        assert torch.equal(run_conv(x[i:i + 1], w, b, 1, None, 1, r[i:i + 1], split=True)[0], full[i])


def test_split_elementwise_ops():
    """pack / unpack / fuse_sum / upsample2x / maxpool / avgpool carry hi + lo pairs"""
    dev = 'cuda'
    g = torch.Generator().manual_seed(1)
    img = torch.randn(2, 3, 224, 224, generator=g)
    o = engine.alloc_act(16, 2, 224, 224, dev, split=True)
    imgd = img.to(dev)
    L.run_op(L.PackImage(imgd.data_ptr(), o.desc()), stream())
    sync_or_die()
    got = engine.from_planar(o).cpu()
    assert torch.equal(got[:, :3].double(), split16(img)) and got[:, 3:].abs().sum() == 0
    outf = torch.zeros(2, 3, 224, 224, device=dev)
    L.run_op(L.Unpack(o.desc(), outf.data_ptr(), 3), stream())
    sync_or_die()
    assert torch.equal(outf.cpu().double(), split16(img))
    # im2col stem packing
    col = engine.alloc_act(32, 2, 112, 112, dev, split=True)
    L.run_op(L.PackImage(imgd.data_ptr(), col.desc(), 1, 0), stream())
    sync_or_die()
    cols = F.unfold(img, 3, padding=1, stride=2).view(2, 3, 9, 112, 112).permute(0, 2, 1, 3, 4).reshape(2, 27, 112, 112)
    gc = engine.from_planar(col).cpu()
    assert torch.equal(gc[:, :27].double(), split16(cols)) and gc[:, 27:].abs().sum() == 0
    # fuse
    xs = [torch.randn(2, 32, 16 >> s, 16 >> s, generator=g) for s in range(3)]
    acts = [engine.to_planar(x.to(dev), split=True) for x in xs]
    out = engine.alloc_act(32, 2, 16, 16, dev, split=True)
    d = L.FuseSum()
    d.out = out.desc()
    for i, a in enumerate(acts):
        d.in_[i] = a.desc()
        d.shift[i] = i
    d.n_in, d.relu = 3, 1
    L.run_op(d, stream())
    sync_or_die()
    xv = [split16(x) for x in xs]
    ref = F.relu(xv[0] + F.interpolate(xv[1], scale_factor=2, mode='nearest') + F.interpolate(xv[2], scale_factor=4, mode='nearest'))
    assert rel_err(engine.from_planar(out).cpu().numpy(), ref.numpy()) < 1e-6
    up = engine.alloc_act(32, 2, 32, 32, dev, split=True)
    L.run_op(L.Upsample2x(acts[0].desc(), up.desc()), stream())
    sync_or_die()
    ref = F.interpolate(xv[0], scale_factor=2, mode='bilinear', align_corners=True)
    assert rel_err(engine.from_planar(up).cpu().numpy(), ref.numpy()) < 1e-6
    for H in (16, 15):
        xm = torch.randn(1, 16, H, H, generator=g)
        am = engine.to_planar(xm.to(dev), split=True)
        Ho = (H - 1) // 2 + 1
        om = engine.alloc_act(16, 1, Ho, Ho, dev, split=True)
        L.run_op(L.MaxPool(am.desc(), om.desc()), stream())
        sync_or_die()
        assert rel_err(engine.from_planar(om).cpu().numpy(), F.max_pool2d(split16(xm), 3, 2, 1).numpy()) < 1e-6
    mat = torch.zeros(2, 40, device=dev)
    L.run_op(L.AvgPool(acts[0].desc(), mat.data_ptr() + 4 * 5, 40), stream())
    sync_or_die()
    assert rel_err(mat[:, 5:37].cpu().numpy(), xv[0].mean(dim=(2, 3)).numpy()) < 1e-6
    # mixing precision modes inside one op is an error, not a silent fp16 result
    bad = L.Upsample2x(acts[0].desc(), engine.alloc_act(32, 2, 32, 32, dev).desc())
    with pytest.raises(L.PocoError):
        L.run_op(bad, stream())


def test_pare_head_split():
    meta, gold, sd = load_preset('pare_w32')
    g = torch.Generator().manual_seed(4)
    N, H = 3, 56
    part = F.relu(torch.randn(N, 128, H, H, generator=g))
    smpl = F.relu(torch.randn(N, 128, H, H, generator=g))
    dev = 'cuda'
    pa, sa = engine.to_planar(part.to(dev), split=True), engine.to_planar(smpl.to(dev), split=True)
    f = lambda *s: torch.zeros(*s, device=dev)
    segm, uf, p6, rot, shape, cam = f(N, 25, H, H), f(N, 3072), f(N, 24, 6), f(N, 24, 3, 3), f(N, 10), f(N, 3)
    scratch = f(int(L.lib().poco_pare_scratch_floats(N, H, H)))
    w = {k: sd['head.' + k].to(dev).contiguous() for k in (
        'keypoint_final_layer.weight', 'keypoint_final_layer.bias', 'smpl_final_layer.weight', 'smpl_final_layer.bias',
        'pose_mlp.weight', 'shape_mlp.weight', 'shape_mlp.bias', 'cam_mlp.weight', 'cam_mlp.bias')}
    d = L.PareHead(pa.desc(), sa.desc(), *(w[k].data_ptr() for k in w), segm.data_ptr(), uf.data_ptr(), p6.data_ptr(),
                   rot.data_ptr(), shape.data_ptr(), cam.data_ptr(), scratch.data_ptr())
    L.run_op(d, stream())
    sync_or_die()
    s = O._SD(sd, 'head.')
    part, smpl = split16(part).float(), split16(smpl).float()
    with torch.no_grad():
        segm_ref = O.conv(part, s, 'keypoint_final_layer')
        att = F.softmax(segm_ref[:, 1:].reshape(N, 24, -1), -1)
        pl = torch.matmul(att, smpl.reshape(N, 128, -1).transpose(2, 1)).transpose(2, 1)
        pose6 = torch.einsum('bcj,ocj->boj', pl, s['pose_mlp.weight'][0, :, :, :, 0, 0]).transpose(2, 1)
    for got, ref, name in ((segm, segm_ref, 'segm'), (uf, pl.reshape(N, -1), 'uncert_feat'), (p6, pose6, 'pose6d')):
        assert rel_err(got.cpu().numpy(), ref.numpy()) < 2e-5, name


@pytest.mark.parametrize('preset', PRESETS)
def test_parity_mode_meets_the_north_star_gate(preset):
    """POCO(precision='split').forward vs the reference-generated goldens: <= 1e-3 on all four gated outputs"""
    meta, gold, _ = load_preset(preset)
    m = build_model(preset, 'cuda', precision='split')
    batch = synthetic_batch(preset, 'cuda')
    with torch.no_grad():
        out = m(batch)
        again = m(batch)                # second call = CUDA-graph capture, third = replay
        third = m(batch)
    sync_or_die(120)
    errs = {k: rel_err(out[k].cpu().numpy(), gold[k]) for k in GATED}
    print(preset, 'split', errs)
    for k in GATED:
        assert errs[k] < PARITY_TOL, (k, errs)
        assert torch.equal(out[k], again[k]) and torch.equal(out[k], third[k]), k
    for k in ('uncert_feat', 'body_feat2', 'pred_pose6d', 'pred_pose_6d'):
        if k in gold:
            assert rel_err(out[k].cpu().numpy(), gold[k]) < PARITY_TOL, k
    if 'pred_segm_mask_sub' in gold:
        st = int(gold['segm_stride'])
        assert rel_err(out['pred_segm_mask'][:, :, ::st, ::st].cpu().numpy(), gold['pred_segm_mask_sub']) < PARITY_TOL
