"""-m gpu: every CUDA op through the C ABI against the oracle arithmetic (torch fp32 on CPU).
Tolerances: conv outputs are stored in fp16 -> 2^-10 relative to the tensor max (one rounding) plus
fp32 accumulation-order noise; fp32 ops -> 1e-5."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import load_preset, rel_err
from gpu_util import conv_reference, run_conv, stream, sync_or_die
from oracle import poco_oracle as O
from poco_b200 import _lib as L
from poco_b200 import engine

pytestmark = pytest.mark.gpu
CONV_TOL = 1.5e-3

# (Cin, Cout, k, stride, H, N, residual, relu)
CONV_CASES = [
    # linear-halo mode (3x3 s1 / 1x1 s1), weights resident
    (32, 32, 3, 1, 56, 2, True, 1), (64, 64, 3, 1, 28, 2, True, 1), (16, 32, 3, 1, 8, 1, False, 0),
    (48, 48, 3, 1, 56, 1, True, 1), (64, 256, 1, 1, 56, 1, True, 1), (256, 64, 1, 1, 56, 1, False, 1),
    (256, 32, 1, 1, 7, 3, False, 0),
    # linear-halo mode, weights streamed per K chunk
    (128, 128, 3, 1, 14, 3, True, 1), (256, 256, 3, 1, 7, 5, True, 1), (480, 256, 3, 1, 28, 1, False, 1),
    (256, 256, 3, 1, 56, 1, False, 1),
    # several N blocks
    (384, 384, 3, 1, 7, 2, True, 1), (1024, 2048, 1, 1, 7, 2, False, 1), (96, 480, 1, 1, 14, 1, False, 0),
    # gather mode (stride 2, 7x7)
    (3, 64, 3, 2, 224, 1, False, 1), (64, 64, 3, 2, 112, 1, False, 1), (32, 64, 3, 2, 56, 2, False, 1),
    (256, 256, 3, 2, 14, 2, True, 2), (3, 64, 7, 2, 224, 1, False, 1), (256, 512, 1, 2, 56, 1, False, 0),
    (128, 256, 3, 2, 14, 3, True, 1),
    # more stride-2 3x3 shapes of HRNet (fuse layers, stem, transitions)
    (32, 32, 3, 2, 56, 3, False, 1), (64, 128, 3, 2, 28, 5, True, 1), (32, 128, 3, 2, 28, 2, True, 0),
    (256, 64, 3, 2, 56, 1, False, 1), (48, 96, 3, 2, 28, 2, True, 1), (32, 32, 3, 2, 28, 3, False, 1),
    (64, 256, 3, 2, 14, 2, True, 1), (16, 64, 3, 2, 224, 2, False, 1),
]


def _case_tensors(cin, cout, k, stride, H, N, res, seed=0):
    g = torch.Generator().manual_seed(seed + cin * 7 + cout)
    x = torch.randn(N, cin, H, H, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = 0.1 * torch.randn(cout, generator=g)
    pad = k // 2
    Ho = (H + 2 * pad - k) // stride + 1
    r = torch.randn(N, cout, Ho, Ho, generator=g) if res else None
    return x, w, b, r


@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'c%d-%d_k%d_s%d_h%d_n%d' % c[:6])
def test_conv_tcgen05(case):
    cin, cout, k, stride, H, N, res, relu = case
    x, w, b, r = _case_tensors(cin, cout, k, stride, H, N, res)
    ref = conv_reference(x, w, b, stride, None, relu, r)
    out = run_conv(x, w, b, stride, None, relu, r, impl=0)
    assert rel_err(out.numpy(), ref.numpy()) < CONV_TOL


@pytest.mark.parametrize('case', [CONV_CASES[0], CONV_CASES[7], CONV_CASES[14], CONV_CASES[17]],
                         ids=lambda c: 'c%d-%d_k%d_s%d_h%d_n%d' % c[:6])
def test_conv_debug_kernel(case):
    cin, cout, k, stride, H, N, res, relu = case
    x, w, b, r = _case_tensors(cin, cout, k, stride, H, N, res)
    ref = conv_reference(x, w, b, stride, None, relu, r)
    out = run_conv(x, w, b, stride, None, relu, r, impl=1)
    assert rel_err(out.numpy(), ref.numpy()) < CONV_TOL


def test_conv_identity_and_tap_shift():
    """1x1 identity weights return the input; a single hot tap of a 3x3 returns the shifted input"""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 12, 12, generator=g).half().float()
    w1 = torch.eye(32).view(32, 32, 1, 1)
    assert torch.equal(run_conv(x, w1, torch.zeros(32), relu=0), x)
    for t in range(9):
        w = torch.zeros(32, 32, 3, 3)
        w[:, :, t // 3, t % 3] = torch.eye(32)
        out = run_conv(x, w, torch.zeros(32), relu=0)
        ref = F.conv2d(x, w, padding=1)
        assert torch.equal(out, ref), f'tap {t}'


def test_conv_batch_invariance():
    """per-crop math does not depend on the batch it is in (needed for bitwise shard == single-GPU)"""
    x, w, b, _ = _case_tensors(64, 64, 3, 1, 28, 4, False)
    full = run_conv(x, w, b)
    for i in range(4):
        assert torch.equal(run_conv(x[i:i + 1], w, b)[0], full[i])


def test_pack_unpack_fuse_upsample_pool():
    dev = 'cuda'
    g = torch.Generator().manual_seed(1)
    img = torch.randn(2, 3, 224, 224, generator=g)
    o = engine.alloc_act(16, 2, 224, 224, dev)
    imgd = img.to(dev)
    L.run_op(L.PackImage(imgd.data_ptr(), o.desc()), stream())
    sync_or_die()
    got = engine.from_planar(o).cpu()
    assert torch.equal(got[:, :3], img.half().float()) and got[:, 3:].abs().sum() == 0
    # unpack
    outf = torch.zeros(2, 3, 224, 224, device=dev)
    L.run_op(L.Unpack(o.desc(), outf.data_ptr(), 3), stream())
    sync_or_die()
    assert torch.equal(outf.cpu(), img.half().float())
    # fuse: x0 + up2(x1) + up4(x2), relu
    xs = [torch.randn(2, 32, 16 >> s, 16 >> s, generator=g).half().float() for s in range(3)]
    acts = [engine.to_planar(x.to(dev)) for x in xs]
    out = engine.alloc_act(32, 2, 16, 16, dev)
    d = L.FuseSum()
    d.out = out.desc()
    for i, a in enumerate(acts):
        d.in_[i] = a.desc()
        d.shift[i] = i
    d.n_in, d.relu = 3, 1
    L.run_op(d, stream())
    sync_or_die()
    ref = F.relu(xs[0] + F.interpolate(xs[1], scale_factor=2, mode='nearest') + F.interpolate(xs[2], scale_factor=4, mode='nearest'))
    assert rel_err(engine.from_planar(out).cpu().numpy(), ref.numpy()) < 1e-3
    # bilinear x2 align_corners
    x = xs[0]
    up = engine.alloc_act(32, 2, 32, 32, dev)
    L.run_op(L.Upsample2x(acts[0].desc(), up.desc()), stream())
    sync_or_die()
    ref = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
    assert rel_err(engine.from_planar(up).cpu().numpy(), ref.numpy()) < 1e-3
    # maxpool 3x3 s2 p1 (odd and even sizes)
    for H in (16, 15):
        xm = torch.randn(1, 16, H, H, generator=g).half().float()
        am = engine.to_planar(xm.to(dev))
        Ho = (H - 1) // 2 + 1
        om = engine.alloc_act(16, 1, Ho, Ho, dev)
        L.run_op(L.MaxPool(am.desc(), om.desc()), stream())
        sync_or_die()
        assert torch.equal(engine.from_planar(om).cpu(), F.max_pool2d(xm, 3, 2, 1))
    # global average pool into a strided matrix
    mat = torch.zeros(2, 40, device=dev)
    L.run_op(L.AvgPool(acts[0].desc(), mat.data_ptr() + 4 * 5, 40), stream())
    sync_or_die()
    assert rel_err(mat[:, 5:37].cpu().numpy(), xs[0].mean(dim=(2, 3)).numpy()) < 1e-5
    assert mat[:, :5].abs().sum() == 0 and mat[:, 37:].abs().sum() == 0


@pytest.mark.parametrize('M,I,O,act', [(4, 2208, 1024, 0), (256, 1024, 144, 0), (7, 3288, 512, 1), (3, 432, 24, 2), (33, 65, 67, 1)])
def test_linear(M, I, O, act):
    g = torch.Generator().manual_seed(2)
    x, w, b, r = torch.randn(M, I, generator=g), torch.randn(O, I, generator=g) / I ** 0.5, torch.randn(O, generator=g), torch.randn(M, O, generator=g)
    xd, wd, bd, rd = (t.cuda() for t in (x, w, b, r))
    y = torch.zeros(M, O + 3, device='cuda')
    L.run_op(L.Linear(xd.data_ptr(), I, wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), O, y.data_ptr() + 4, O + 3, M, I, O, act), stream())
    sync_or_die()
    # the same problem with a split-K workspace (several CTAs per output tile + reduce kernel)
    y2 = torch.zeros(M, O + 3, device='cuda')
    scratch = torch.empty(8 * M * O, device='cuda')
    L.run_op(L.Linear(xd.data_ptr(), I, wd.data_ptr(), bd.data_ptr(), rd.data_ptr(), O, y2.data_ptr() + 4, O + 3, M, I, O, act,
                      scratch.data_ptr(), scratch.numel()), stream())
    sync_or_die()
    assert rel_err(y2[:, 1:O + 1].cpu().numpy(), y[:, 1:O + 1].cpu().numpy()) < 1e-5
    assert y2[:, 0].abs().sum() == 0 and y2[:, O + 1:].abs().sum() == 0
    ref = x.double() @ w.double().t() + b.double()
    ref = torch.sigmoid(ref) if act == 1 else (F.softplus(ref) if act == 2 else ref)
    ref = ref + r.double()
    assert rel_err(y[:, 1:O + 1].cpu().numpy(), ref.numpy()) < 1e-5
    assert y[:, 0].abs().sum() == 0 and y[:, O + 1:].abs().sum() == 0


def test_rot6d_and_copy2d():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 160, generator=g)
    xd = x.cuda()
    out = torch.zeros(5 * 24, 3, 3, device='cuda')
    L.run_op(L.Rot6d(xd.data_ptr() + 4 * 10, 160, 24, 120, out.data_ptr()), stream())
    sync_or_die()
    ref = O.rot6d_to_rotmat(x[:, 10:154].reshape(-1, 6))
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5
    dst = torch.zeros(5, 20, device='cuda')
    src = torch.arange(7.0).view(1, 7).cuda()
    L.run_op(L.Copy2d(src.data_ptr(), 7, dst.data_ptr() + 4 * 3, 20, 5, 7, 1), stream())
    sync_or_die()
    assert torch.equal(dst[:, 3:10].cpu(), torch.arange(7.0).expand(5, 7))


@pytest.mark.parametrize('H', [56, 7])
def test_pare_head(H):
    """fused part-attention head vs the oracle restatement of pare_head.py:754-826, :896-906"""
    meta, gold, sd = load_preset('pare_w32')
    g = torch.Generator().manual_seed(4)
    N = 3
    part = F.relu(torch.randn(N, 128, H, H, generator=g)).half().float()
    smpl = F.relu(torch.randn(N, 128, H, H, generator=g)).half().float()
    dev = 'cuda'
    pa, sa = engine.to_planar(part.to(dev)), engine.to_planar(smpl.to(dev))
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
    # oracle: same math from the two branch outputs on
    s = O._SD(sd, 'head.')
    with torch.no_grad():
        segm_ref = O.conv(part, s, 'keypoint_final_layer')
        att = F.softmax(segm_ref[:, 1:].reshape(N, 24, -1), -1)
        pl = torch.matmul(att, smpl.reshape(N, 128, -1).transpose(2, 1)).transpose(2, 1)
        cs = torch.matmul(att, O.conv(smpl, s, 'smpl_final_layer').reshape(N, 64, -1).transpose(2, 1)).transpose(2, 1)
        pose6 = torch.einsum('bcj,ocj->boj', pl, s['pose_mlp.weight'][0, :, :, :, 0, 0]).transpose(2, 1)
        shape_ref = O.linear(cs.flatten(1), s, 'shape_mlp')
        cam_ref = O.linear(cs.flatten(1), s, 'cam_mlp')
        rot_ref = O.rot6d_to_rotmat(pose6).reshape(N, 24, 3, 3)
    for got, ref, name in ((segm, segm_ref, 'segm'), (uf, pl.reshape(N, -1), 'uncert_feat'), (p6, pose6, 'pose6d'),
                           (shape, shape_ref, 'shape'), (cam, cam_ref, 'cam'), (rot, rot_ref, 'rotmat')):
        assert rel_err(got.cpu().numpy(), ref.numpy()) < 2e-5, name


@pytest.mark.parametrize('preset', ['pare_w32', 'cliff_w32'])
def test_realnvp_against_reference_goldens(preset):
    from common import build_model
    meta, gold, sd = load_preset(preset)
    m = build_model(preset, 'cuda')
    rows = gold['flow_x'].shape[0]
    ctx = torch.repeat_interleave(torch.from_numpy(gold['flow_ctx']), rows // meta['test_b'], 0).cuda()
    x, z = torch.from_numpy(gold['flow_x']).cuda(), torch.from_numpy(gold['flow_z']).cuda()
    lp = m.flow_log_prob(x, ctx)
    zb, ld = m.flow_backward(x, ctx)
    fx = m.flow_forward(z, ctx)
    sync_or_die()
    assert rel_err(lp.cpu().numpy(), gold['flow_log_prob']) < 2e-5
    assert rel_err(zb.cpu().numpy(), gold['flow_backward_z']) < 2e-5
    assert rel_err(ld.cpu().numpy(), gold['flow_logdet']) < 2e-5
    assert rel_err(fx.cpu().numpy(), gold['flow_forward_x']) < 2e-5
    # round trip and ragged row count (R not a multiple of the CTA row block)
    assert rel_err(m.flow_forward(zb[:13], ctx[:13]).cpu().numpy(), gold['flow_x'][:13]) < 1e-4
    # cond_layer
    uf = torch.from_numpy(gold['uncert_feat']).cuda()
    assert rel_err(m.flow_context(uf).cpu().numpy(), gold['flow_ctx']) < 2e-5
    # one context row per crop (the 24 joints of a crop share it, nf_head.py:85-101): same numbers as the expanded call
    per = rows // meta['test_b']
    ctx1 = torch.from_numpy(gold['flow_ctx']).cuda()
    lp1 = m.flow_log_prob(x, ctx1, rows_per_ctx=per)
    fx1 = m.flow_forward(z, ctx1, rows_per_ctx=per)
    sync_or_die()
    assert torch.equal(lp1, lp) and torch.equal(fx1, fx)
    # the un-hoisted kernel path (ctx_part = NULL in the C ABI) still agrees
    from poco_b200 import _lib as L2
    params, nl = m._flow_params(x.device)[:2]
    out = torch.empty(rows, device='cuda')
    L2.run_op(L2.RealNVP(x.data_ptr(), ctx.data_ptr(), params.data_ptr(), out.data_ptr(), None, None,
                         rows, x.shape[1], ctx.shape[1], 64, nl, 0, None, 1, 0), stream())
    sync_or_die()
    assert rel_err(out.cpu().numpy(), gold['flow_log_prob']) < 2e-5


# ------------------------------------------------------------------------------------------------
# conv chains: the four BasicBlocks of an HRNet branch as one persistent launch (poco_conv_chain)
# ------------------------------------------------------------------------------------------------
def _branch_ops(ch, H, N, chained, max_ctas=0, seed=0, groups=1):
    """ops + buffers of one HRNet branch (4 BasicBlocks, hrnet.py:42-58) built by the plan builder"""
    from poco_b200 import arch
    g = torch.Generator().manual_seed(seed + ch)
    sd = {}
    for k in range(4):
        for c, bn, gamma in (('conv1', 'bn1', 1.0), ('conv2', 'bn2', 0.4)):
            sd[f'br.{k}.{c}.weight'] = torch.randn(ch, ch, 3, 3, generator=g) * (2.0 / (9 * ch)) ** 0.5
            sd[f'br.{k}.{bn}.weight'] = gamma * (0.8 + 0.4 * torch.rand(ch, generator=g))
            sd[f'br.{k}.{bn}.bias'] = 0.1 * torch.randn(ch, generator=g)
            sd[f'br.{k}.{bn}.running_mean'] = 0.1 * torch.randn(ch, generator=g)
            sd[f'br.{k}.{bn}.running_var'] = 0.8 + 0.4 * torch.rand(ch, generator=g)
    b = engine.PlanBuilder(sd, N, 'cuda')
    b.use_chains = chained
    if max_ctas:
        b.shares, b.lane = [max_ctas], 0
    x0 = torch.randn(N, ch, H, H, generator=g)
    x = engine.to_planar(x0.cuda())
    b.keep.append(x.buf)
    xin = x
    b.begin_chain()
    for k in range(4):
        x = arch.basic_block(b, x, f'br.{k}', ch, ch)
    b.end_chain(groups=groups)
    return b, xin, x, x0, sd


@pytest.mark.parametrize('ch,H,N,max_ctas,groups', [(32, 56, 12, 0, 1), (32, 56, 5, 7, 1), (64, 28, 20, 0, 1),
                                                    (128, 14, 40, 0, 1), (256, 7, 64, 0, 1), (128, 14, 9, 3, 1),
                                                    (48, 56, 3, 0, 1), (32, 56, 11, 0, 3), (256, 7, 13, 0, 4),
                                                    (64, 28, 9, 5, 2)],
                         ids=lambda v: str(v))
def test_conv_chain_matches_separate_launches(ch, H, N, max_ctas, groups):
    """one chained launch (or one per crop range) == eight separate conv launches (four fused BasicBlock launches for
    32 channels) up to fp16 rounding flips, and == the fp32 oracle arithmetic"""
    outs = []
    for chained in (False, True):
        b, xin, xout, x0, sd = _branch_ops(ch, H, N, chained, max_ctas, groups=groups)
        kinds = [op.kind for op in b.ops]
        # (unchained 32-channel blocks run as one fused poco_basic_block launch each)
        separate = [L.OP_BASIC_BLOCK] * 4 if L.lib().poco_basic_block_supported(ch, H, H) else [L.OP_CONV] * 8
        assert kinds == ([L.OP_CONV_CHAIN] * groups if chained else separate)
        for rep in range(3 if chained else 1):          # replays re-zero the tile flags
            engine.act_view(xin)[:, :, 1:H + 1, 1:H + 1, :] = \
                x0.cuda().half().view(N, ch // 8, 8, H, H).permute(1, 0, 3, 4, 2)
            for op in b.ops:
                L.run_op(op, stream())
            sync_or_die(30)
            outs.append(engine.from_planar(xout).cpu())
        halo = engine.act_view(xout)
        assert float(halo[:, :, 0].abs().sum() + halo[:, :, -1].abs().sum() + halo[:, :, :, 0].abs().sum() +
                     halo[:, :, :, -1].abs().sum()) == 0.0
    for o in outs[2:]:
        assert torch.equal(o, outs[1])                  # replays of the chain are deterministic
    # the separate launches may pick another configuration than the chain (two-CTA "half" mode for N <= 64, N blocks of
    # 128 columns with two tiles per streamed weight stage for the wide layers): other K chunking, so the fp32
    # accumulation order differs and a few fp16 roundings flip per layer
    assert rel_err(outs[1].numpy(), outs[0].numpy()) < 4e-3
    # oracle arithmetic with the engine's roundings (fp16 activations and folded weights, fp32 accumulate)
    x = x0.half().float()
    for k in range(4):
        y = x
        for c, bn, relu, res in (('conv1', 'bn1', True, False), ('conv2', 'bn2', True, True)):
            bnp = tuple(sd[f'br.{k}.{bn}{s}'] for s in ('.weight', '.bias', '.running_mean', '.running_var'))
            wf, bf = engine.fold_bn(sd[f'br.{k}.{c}.weight'], None, bnp)
            y = F.conv2d(y, wf.half().float(), bf, padding=1)
            if res:
                y = y + x
            y = F.relu(y).half().float()
        x = y
    assert rel_err(outs[0].numpy(), x.numpy()) < 4e-3


@pytest.mark.parametrize('C,H,N', [(24, 28, 3), (48, 14, 2), (64, 56, 2), (256, 7, 5)])
def test_fuse_sum_and_upsample_odd_geometries(C, H, N):
    """row-mapped element-wise kernels: widths that are not powers of two, plane counts 3 / 6 / 8 / 32"""
    dev = 'cuda'
    g = torch.Generator().manual_seed(C + H)
    xs = [torch.randn(N, C, H, H, generator=g).half().float(), torch.randn(N, C, H // 2, H // 2, generator=g).half().float()] \
        if H % 2 == 0 else [torch.randn(N, C, H, H, generator=g).half().float()]
    acts = [engine.to_planar(x.to(dev)) for x in xs]
    out = engine.alloc_act(C, N, H, H, dev)
    d = L.FuseSum()
    d.out = out.desc()
    for i, a in enumerate(acts):
        d.in_[i] = a.desc()
        d.shift[i] = i
    d.n_in, d.relu = len(acts), 0
    L.run_op(d, stream())
    sync_or_die()
    ref = xs[0] + (F.interpolate(xs[1], scale_factor=2, mode='nearest') if len(xs) > 1 else 0)
    assert rel_err(engine.from_planar(out).cpu().numpy(), ref.numpy()) < 1e-3
    hv = engine.act_view(out)
    assert float(hv[:, :, 0].abs().sum() + hv[:, :, -1].abs().sum() + hv[:, :, :, 0].abs().sum() + hv[:, :, :, -1].abs().sum()) == 0.0
    if C % 16 == 0:
        up = engine.alloc_act(C, N, 2 * H, 2 * H, dev)
        L.run_op(L.Upsample2x(acts[0].desc(), up.desc()), stream())
        sync_or_die()
        ref = F.interpolate(xs[0], scale_factor=2, mode='bilinear', align_corners=True)
        assert rel_err(engine.from_planar(up).cpu().numpy(), ref.numpy()) < 1e-3
        hv = engine.act_view(up)
        assert float(hv[:, :, 0].abs().sum() + hv[:, :, -1].abs().sum() + hv[:, :, :, 0].abs().sum() + hv[:, :, :, -1].abs().sum()) == 0.0


def test_stem_im2col_conv_matches_strided_conv():
    """pack_image(im2col=1) + 1x1 conv (K = 32) == conv3x3(3 -> 64, stride 2, pad 1) + BN + ReLU of the HRNet stem"""
    from poco_b200 import arch
    g = torch.Generator().manual_seed(11)
    N, H = 3, 64
    img = torch.randn(N, 3, H, H, generator=g)
    sd = {'c.weight': torch.randn(64, 3, 3, 3, generator=g) * 0.2, 'b.weight': 0.8 + 0.4 * torch.rand(64, generator=g),
          'b.bias': 0.1 * torch.randn(64, generator=g), 'b.running_mean': 0.1 * torch.randn(64, generator=g),
          'b.running_var': 0.8 + 0.4 * torch.rand(64, generator=g)}
    b = engine.PlanBuilder(sd, N, 'cuda')
    imgd = img.cuda()
    out = b.stem_conv(imgd, H, H, 'c', 'b', 64)
    for op in b.ops:
        L.run_op(op, stream())
    sync_or_die()
    wf, bf = engine.fold_bn(sd['c.weight'], None, tuple(sd['b' + s] for s in ('.weight', '.bias', '.running_mean', '.running_var')))
    ref = F.relu(F.conv2d(img.half().float(), wf.half().float(), bf, stride=2, padding=1))
    assert rel_err(engine.from_planar(out).cpu().numpy(), ref.numpy()) < CONV_TOL


# dx-in-N mode (poco_conv.wfmt = 1): (Cin, Cout, H, N, residual, relu)
DXN_CASES = [(32, 32, 56, 2, True, 1), (32, 32, 56, 1, False, 0), (64, 64, 28, 3, True, 1), (64, 64, 56, 1, False, 1),
             (256, 32, 56, 1, False, 1), (16, 32, 8, 1, False, 0), (128, 64, 14, 5, True, 1), (32, 64, 7, 9, False, 1),
             (64, 32, 126, 1, False, 1)]


@pytest.mark.parametrize('case', DXN_CASES, ids=lambda c: 'c%d-%d_h%d_n%d' % c[:4])
def test_conv_dx_in_n(case):
    cin, cout, H, N, res, relu = case
    x, w, b, r = _case_tensors(cin, cout, 3, 1, H, N, res, seed=3)
    ref = conv_reference(x, w, b, 1, None, relu, r)
    out = run_conv(x, w, b, 1, None, relu, r, impl=0, dxn=True)
    assert rel_err(out.numpy(), ref.numpy()) < CONV_TOL


def test_conv_dx_in_n_tap_shift():
    """a single hot tap returns the shifted input exactly (checks the shuffle / exchange directions and the
    126-pixel tile stride at warp and tile boundaries)"""
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 32, 20, 20, generator=g).half().float()
    for t in range(9):
        w = torch.zeros(32, 32, 3, 3)
        w[:, :, t // 3, t % 3] = torch.eye(32)
        out = run_conv(x, w, torch.zeros(32), relu=0, dxn=True)
        assert torch.equal(out, F.conv2d(x, w, padding=1)), f'tap {t}'


def test_crop_normalize_matches_reference_bit_for_bit():
    """poco_crop against the reference-generated golden crops and, on fresh random detections,
    against the oracle: integer warp arithmetic -> bit-exact"""
    import os

    from oracle import crop_oracle as C
    from poco_b200 import crop_batch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'crop_golden.npz'))
    frame = torch.from_numpy(g['frame']).cuda()
    out = crop_batch(frame, torch.from_numpy(g['boxes']), scale=float(g['scale']))
    sync_or_die()
    assert np.array_equal(out['img'].cpu().numpy(), g['img'])
    assert np.array_equal(out['bbox_info'].cpu().numpy(), g['bbox_info'])
    assert np.array_equal(out['focal_length'].cpu().numpy(), g['focal_length'])
    # fresh detections on a larger frame (1080p), several far outside it; empty crops are all (0 - mean) / std
    fr = C.synthetic_frame(5, 1080, 1920)
    bx = C.synthetic_boxes(7, 24, 1080, 1920)
    bx[3] = [-500, -500, 100, 100]
    ref = C.crop_batch(fr, bx, 1.1)
    got = crop_batch(torch.from_numpy(fr).cuda(), torch.from_numpy(bx.astype(np.float32)), scale=1.1)
    sync_or_die()
    for k in ('img', 'bbox_info', 'focal_length', 'scale', 'center', 'orig_shape'):
        assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
    assert np.unique(ref['img'][3][0]).size == 1
    # the product path has no CPU route
    with pytest.raises(Exception):
        crop_batch(torch.from_numpy(fr), torch.from_numpy(bx.astype(np.float32)))


def test_uncert_post_matches_reference():
    """poco_uncert_post against the reference-generated golden: the element-wise parts bit for bit,
    the PARE row mean to one fp32 rounding"""
    import os

    from poco_b200 import uncert_post
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'uncert_golden.npz'))
    var = torch.from_numpy(g['var']).cuda()
    for bb, name in (('cliff', 'hrnet_w48_cls-cliff'), ('pare', 'hrnet_w32-pare')):
        for kin in (0, 1):
            tag = f'{bb}_{kin}'
            p, t, gl = uncert_post(var, name, kinematic=bool(kin))
            sync_or_die()
            assert np.array_equal(p.cpu().numpy(), g['prepared_' + tag])
            assert np.array_equal(t.cpu().numpy(), g['thresholded_' + tag])
            assert np.allclose(gl.cpu().numpy(), g['global_' + tag], rtol=2e-7, atol=0)
    conf, _, _ = uncert_post(var, 'hrnet_w32-pare', return_conf=True)
    assert np.array_equal(conf.cpu().numpy(), 1 - g['var'])
    assert np.array_equal(var.cpu().numpy(), g['var'])              # the input is never modified


# ------------------------------------------------------------------------------------------------
# space-to-depth plumbing of the stride-2 convs (poco_conv.out_s2d / in_s2d)
# ------------------------------------------------------------------------------------------------
def _s2d_plan(cin, cmid, cout, H, N, use_s2d, split, mode, seed=0):
    """conv3x3 s1 (cin -> cmid, writes the phase-split copy) -> conv3x3 s2 (cmid -> cout, + residual) -> conv1x1 s2"""
    import os
    g = torch.Generator().manual_seed(seed + cin + cout)
    sd = {}
    for name, ci, co, k in (('a', cin, cmid, 3), ('b', cmid, cout, 3), ('c', cmid, cout, 1)):
        sd[f'{name}.weight'] = torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5
        sd[f'{name}bn.weight'] = 0.8 + 0.4 * torch.rand(co, generator=g)
        sd[f'{name}bn.bias'] = 0.1 * torch.randn(co, generator=g)
        sd[f'{name}bn.running_mean'] = 0.1 * torch.randn(co, generator=g)
        sd[f'{name}bn.running_var'] = 0.8 + 0.4 * torch.rand(co, generator=g)
    os.environ['POCO_B200_S2D'] = '2' if use_s2d else '0'
    try:
        b = engine.PlanBuilder(sd, N, 'cuda', split=split)
        x0 = torch.randn(N, cin, H, H, generator=g)
        r0 = torch.randn(N, cout, H // 2, H // 2, generator=g)
        x = engine.to_planar(x0.cuda(), split=split)
        r = engine.to_planar(r0.cuda(), split=split)
        b.keep += [x.buf, r.buf]
        y = b.conv_bn(x, 'a', 'abn', cin, cmid, 3, s2d=mode)
        z = b.conv_bn(y, 'b', 'bbn', cmid, cout, 3, 2, relu=True, residual=r)
        w = b.conv_bn(y, 'c', 'cbn', cmid, cout, 1, 2, relu=False, pad=0)
    finally:
        os.environ.pop('POCO_B200_S2D')
    return b, y, z, w, x0, r0, sd


@pytest.mark.parametrize('cin,cmid,cout,H,N,split', [(32, 32, 64, 56, 3, False), (64, 64, 128, 28, 2, False),
                                                     (16, 64, 64, 112, 1, False), (128, 128, 256, 14, 3, False),
                                                     (256, 256, 64, 56, 1, False), (32, 32, 64, 56, 2, True),
                                                     (64, 64, 128, 28, 3, True), (128, 128, 256, 14, 2, True)],
                         ids=lambda v: str(v))
def test_stride2_convs_on_phase_split_inputs(cin, cmid, cout, H, N, split):
    """the producing conv writes the phase-split copy from its epilogue ('dual'), the 3x3 / stride 2 conv walks it on
    the halo-run path and the 1x1 / stride 2 conv reads phase block 0: same results as the gather path and as the
    oracle arithmetic"""
    import emu
    res = {}
    for use in (False, True):
        b, y, z, w, x0, r0, sd = _s2d_plan(cin, cmid, cout, H, N, use, split, 'dual')
        kinds = [(op.u.conv.in_s2d, bool(op.u.conv.out_s2d.data), op.u.conv.stride) for op in b.ops]
        assert kinds == ([(0, True, 1), (1, False, 2), (0, False, 1)] if use else [(0, False, 1), (0, False, 2), (0, False, 2)])
        for op in b.ops:
            L.run_op(op, stream())
        sync_or_die(30)
        res[use] = (engine.from_planar(y).cpu(), engine.from_planar(z).cpu(), engine.from_planar(w).cpu())
        if use:     # the phase-split copy holds exactly the values of the normal output
            assert torch.equal(engine.from_planar(y.s2d).cpu(), emu.space_to_depth(res[use][0]))
            for t in (y, y.s2d, z, w):
                for lo in ([False, True] if split else [False]):
                    halo = engine.act_view(t, lo)
                    assert float(halo[:, :, 0].abs().sum() + halo[:, :, -1].abs().sum() + halo[:, :, :, 0].abs().sum() +
                                 halo[:, :, :, -1].abs().sum()) == 0.0, 'kernel wrote into the zero halo'
    tol = 2e-5 if split else CONV_TOL
    assert rel_err(res[True][0].numpy(), res[False][0].numpy()) < tol          # (half-size CTAs without the second output)
    assert rel_err(res[True][1].numpy(), res[False][1].numpy()) < tol          # (other K chunking than the gather)
    assert rel_err(res[True][2].numpy(), res[False][2].numpy()) < tol
    # oracle arithmetic on the engine's operands
    from gpu_util import split16
    rnd = (lambda t: split16(t)) if split else (lambda t: t.half().double())

    def cbr(x, name, stride, pad, relu, resid=None):
        bnp = tuple(sd[f'{name}bn{s_}'] for s_ in ('.weight', '.bias', '.running_mean', '.running_var'))
        wf, bf = engine.fold_bn(sd[f'{name}.weight'], None, bnp)
        v = F.conv2d(x, rnd(wf), bf.double(), stride=stride, padding=pad)
        if resid is not None:
            v = v + resid
        return rnd(F.relu(v) if relu else v)
    yr = cbr(rnd(x0), 'a', 1, 1, True)
    zr = cbr(yr, 'b', 2, 1, True, rnd(r0))
    wr = cbr(yr, 'c', 2, 0, False)
    assert rel_err(res[True][0].numpy(), yr.numpy()) < tol
    assert rel_err(res[True][1].numpy(), zr.numpy()) < (4e-5 if split else 2 * CONV_TOL)
    assert rel_err(res[True][2].numpy(), wr.numpy()) < (4e-5 if split else 2 * CONV_TOL)


def test_phase_split_only_output_leaves_no_normal_tensor():
    b, y, z, w, x0, r0, sd = _s2d_plan(32, 32, 64, 28, 2, True, False, 'only')
    assert y.buf is None and y.s2d is not None and b.ops[0].u.conv.s2d_only == 1
    for op in b.ops[:2]:
        L.run_op(op, stream())
    sync_or_die(30)
    b2, y2, z2, w2, _, _, _ = _s2d_plan(32, 32, 64, 28, 2, False, False, None)
    for op in b2.ops[:2]:
        L.run_op(op, stream())
    sync_or_die(30)
    assert rel_err(engine.from_planar(z).cpu().numpy(), engine.from_planar(z2).cpu().numpy()) < CONV_TOL


@pytest.mark.parametrize('C,H,W,N,max_ctas', [(32, 56, 56, 2, 0), (32, 56, 56, 7, 5), (32, 28, 28, 3, 0), (32, 8, 12, 1, 0), (32, 56, 56, 37, 0),
                                              (32, 20, 61, 2, 3), (64, 28, 28, 3, 0), (64, 28, 28, 9, 5), (64, 14, 20, 2, 0), (64, 56, 56, 2, 0),
                                              (64, 28, 28, 40, 0)])
def test_basic_block_fused_matches_two_convs(C, H, W, N, max_ctas):
    """poco_basic_block (conv1 -> shared memory -> conv2 + input as residual, one launch) against the same block as two
    poco_conv launches: bit-identical (same MMA order, same fp16 rounding of the intermediate; the two-launch block runs
    with a CTA budget so that it takes the one-CTA-per-SM flavour -- the two-CTA flavour walks K in 16-channel chunks,
    another summation order), and against fp32 arithmetic on the fp16-rounded operands (hrnet.py:42-58).  Cases: several units per CTA / one short unit, a CTA
    budget (plan lanes), crops that straddle unit boundaries, the widest row the kernel takes (W + 3 = 64); the
    64-channel flavour (one conv2 tile per unit, single shared-memory buffers)."""
    from gpu_util import run_basic_block
    assert L.lib().poco_basic_block_supported(C, H, W) == 1
    g = torch.Generator().manual_seed(H * 100 + N)
    x = torch.randn(N, C, H, W, generator=g)
    w1, w2 = (torch.randn(C, C, 3, 3, generator=g) * (0.08 if C == 32 else 0.06) for _ in range(2))
    b1, b2 = (torch.randn(C, generator=g) * 0.2 for _ in range(2))
    got = run_basic_block(x, w1, b1, w2, b2, max_ctas)
    mid = run_conv(x, w1, b1, relu=1, max_ctas=148)
    two = run_conv(mid, w2, b2, relu=1, residual=x, max_ctas=148)
    assert torch.equal(got, two), float((got - two).abs().max())
    ref_mid = conv_reference(x, w1, b1, relu=1).half().float()
    ref = conv_reference(ref_mid, w2, b2, relu=1, residual=x)
    assert rel_err(got, ref) < CONV_TOL


@pytest.mark.parametrize('H,W,N,max_ctas', [(56, 56, 3, 0), (28, 28, 5, 3), (8, 12, 2, 0)])
def test_basic_block_fused_writes_the_phase_split_copy(H, W, N, max_ctas):
    """poco_basic_block.out_s2d: the fused 32-channel block writes its output a second time in phase-split form
    (channel ((y & 1) * 2 + (x & 1)) * C + c of pixel (y / 2, x / 2)) for the stride-2 fuse convs (hrnet.py:213-240):
    the normal output is unchanged bit for bit and the copy is exactly its space-to-depth"""
    import emu
    from gpu_util import run_basic_block
    C = 32
    g = torch.Generator().manual_seed(H + N)
    x = torch.randn(N, C, H, W, generator=g)
    w1, w2 = (torch.randn(C, C, 3, 3, generator=g) * 0.08 for _ in range(2))
    b1, b2 = (torch.randn(C, generator=g) * 0.2 for _ in range(2))
    plain = run_basic_block(x, w1, b1, w2, b2, max_ctas)
    out, copy = run_basic_block(x, w1, b1, w2, b2, max_ctas, s2d=True)
    assert torch.equal(out, plain)
    assert torch.equal(copy, emu.space_to_depth(out))


def test_basic_block_rejects_what_it_cannot_fuse():
    lib = L.lib()
    assert lib.poco_basic_block_supported(128, 14, 14) == 0 and lib.poco_basic_block_supported(32, 56, 62) == 0
    assert lib.poco_basic_block_supported(48, 56, 56) == 0 and lib.poco_basic_block_supported(64, 28, 62) == 0
    a = engine.alloc_act(128, 1, 14, 14, 'cuda')
    o = engine.alloc_act(128, 1, 14, 14, 'cuda')
    w = torch.zeros(9 * 16 * 128 * 8, dtype=torch.float16, device='cuda')
    b = torch.zeros(128, device='cuda')
    d = L.BasicBlock(a.desc(), o.desc(), w.data_ptr(), b.data_ptr(), w.data_ptr(), b.data_ptr(), 0, 0)
    with pytest.raises(L.PocoError):
        L.run_op(d, stream())
    a32 = engine.alloc_act(32, 1, 8, 8, 'cuda')
    d = L.BasicBlock(a32.desc(), a32.desc(), w.data_ptr(), b.data_ptr(), w.data_ptr(), b.data_ptr(), 0, 0)
    with pytest.raises(L.PocoError):        # in-place
        L.run_op(d, stream())


@pytest.mark.parametrize('H,W,N,blocks,max_ctas,in_place', [(14, 14, 3, 4, 0, True), (14, 14, 9, 4, 4, True), (14, 14, 2, 1, 0, False),
                                                            (14, 14, 5, 2, 2, False), (7, 7, 4, 4, 0, True), (6, 20, 3, 3, 0, True),
                                                            (14, 14, 200, 4, 0, True)])
def test_branch_resident_matches_separate_convs(H, W, N, blocks, max_ctas, in_place):
    """poco_branch (the BasicBlocks of an HRNet branch in one launch, crop resident in shared memory; hrnet.py:42-58 x4,
    hrnet.py:140-186) against the same blocks as poco_conv launches (same fp16 rounding points: every conv1 output and
    every block output; the streamed poco_conv walks K in another order, so equal to one fp16 rounding per hand-off) and
    against fp32 arithmetic on the fp16-rounded operands.  Cases: several crops per CTA with a CTA budget (plan lanes),
    one block, in place and out of place, a crop of one tile (7x7), a non-square crop that does not fill its two tiles,
    more crops than SMs."""
    from gpu_util import run_branch
    C = 128
    assert L.lib().poco_branch_supported(C, H, W, blocks) == 1
    g = torch.Generator().manual_seed(H * 100 + N)
    x = torch.randn(N, C, H, W, generator=g)
    ws = [torch.randn(C, C, 3, 3, generator=g) * 0.04 for _ in range(2 * blocks)]
    bs = [torch.randn(C, generator=g) * 0.2 for _ in range(2 * blocks)]
    got = run_branch(x, ws, bs, max_ctas, in_place)
    sep, ref = x.half().float(), x.half().float()
    for k in range(blocks):
        if N <= 16:
            mid = run_conv(sep, ws[2 * k], bs[2 * k], relu=1)
            sep = run_conv(mid, ws[2 * k + 1], bs[2 * k + 1], relu=1, residual=sep)
        ref_mid = conv_reference(ref, ws[2 * k], bs[2 * k], relu=1).half().float()
        ref = conv_reference(ref_mid, ws[2 * k + 1], bs[2 * k + 1], relu=1, residual=ref).half().float()
    assert rel_err(got, ref) < 2 * CONV_TOL, rel_err(got, ref)
    if N <= 16:
        assert rel_err(got, sep) < 2 * CONV_TOL, rel_err(got, sep)
    # a crop's result does not depend on the batch it travels in or on the CTA that ran it
    if N >= 3:
        alone = run_branch(x[2:3], ws, bs, 0, in_place)
        assert torch.equal(alone, got[2:3])


def test_branch_rejects_what_it_cannot_run():
    lib = L.lib()
    assert lib.poco_branch_supported(64, 14, 14, 4) == 0 and lib.poco_branch_supported(128, 28, 28, 4) == 0
    assert lib.poco_branch_supported(128, 14, 14, 5) == 0 and lib.poco_branch_supported(128, 14, 14, 0) == 0
    a = engine.alloc_act(128, 1, 28, 28, 'cuda')
    w = torch.zeros(9 * 16 * 128 * 8, dtype=torch.float16, device='cuda')
    b = torch.zeros(128, device='cuda')
    d = L.Branch()
    d.in_, d.out, d.n_blocks = a.desc(), a.desc(), 1
    d.weight[0] = d.weight[1] = w.data_ptr()
    d.bias[0] = d.bias[1] = b.data_ptr()
    with pytest.raises(L.PocoError):        # the crop does not fit
        L.run_op(d, stream())
    a14 = engine.alloc_act(128, 1, 14, 14, 'cuda')
    d.in_, d.out, d.n_blocks = a14.desc(), a14.desc(), 2
    with pytest.raises(L.PocoError):        # null weights of the second block
        L.run_op(d, stream())


@pytest.mark.parametrize('H,W,N,max_ctas', [(56, 56, 2, 0), (56, 56, 5, 7), (28, 28, 3, 0), (8, 12, 1, 0), (56, 56, 19, 0)])
def test_bottleneck_tail_fused_matches_two_convs(H, W, N, max_ctas):
    """poco_bottleneck_tail (3x3 conv -> shared memory -> 1x1 conv + residual, one launch; hrnet.py:88-99) against the
    two poco_conv launches it replaces (same fp16 rounding of the intermediate; the 1x1 conv accumulates its four K steps
    in the same order, the 3x3 in the one-CTA flavour's order) and against fp32 arithmetic on the fp16-rounded operands."""
    from gpu_util import run_bottleneck_tail
    assert L.lib().poco_bottleneck_tail_supported(64, 256, H, W) == 1
    g = torch.Generator().manual_seed(H * 100 + N)
    x = torch.randn(N, 64, H, W, generator=g)
    res = torch.randn(N, 256, H, W, generator=g)
    w2 = torch.randn(64, 64, 3, 3, generator=g) * 0.06
    w3 = torch.randn(256, 64, 1, 1, generator=g) * 0.15
    b2, b3 = torch.randn(64, generator=g) * 0.2, torch.randn(256, generator=g) * 0.2
    got = run_bottleneck_tail(x, w2, b2, w3, b3, res, max_ctas)
    mid = run_conv(x, w2, b2, relu=1, max_ctas=148)
    two = run_conv(mid, w3, b3, relu=1, residual=res, max_ctas=148)
    ref_mid = conv_reference(x, w2, b2, relu=1).half().float()
    ref = conv_reference(ref_mid, w3, b3, relu=1, residual=res)
    assert rel_err(got, ref) < CONV_TOL
    assert rel_err(got, two) < 2e-3 and float((got != two).float().mean()) < 0.01, float((got != two).float().mean())
