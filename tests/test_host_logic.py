"""CPU tests of the host side: checkpoint contract, C-ABI surface, weight preparation, and the op
schedule replayed by the CPU op interpreter (tests/emu.py) against the reference-generated goldens."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import emu
from common import GATED, PRESETS, build_model, load_preset, rel_err, synthetic_batch
from poco_b200 import POCO, PocoError, _lib, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'poco_b200.h')).read()
    declared = set(re.findall(r'\b(poco_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert L.poco_version() >= 100


def test_ctypes_structs_match_header_sizes():
    # poco_act: ptr + i64 + 4*i32 + ptr (lo) = 40 bytes;
    # poco_conv: 2*40 + 3 ptr + i64 + 8*i32 + ptr (residual_lo) + 40 (out_s2d) + 2*i32 = 200
    assert ctypes.sizeof(_lib.Act) == 40
    assert ctypes.sizeof(_lib.Conv) == 200
    assert ctypes.sizeof(_lib.Linear) == 96
    assert _lib.Op.u.offset == 8


def test_ctypes_layout_matches_header_field_by_field(tmp_path):
    """include/poco_b200.h is compiled as plain C99 (the ABI must not need a C++ compiler) and every struct's size and
    every field's offset is compared with the ctypes mirror in poco_b200/_lib.py"""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    pairs = [('poco_act', _lib.Act), ('poco_conv', _lib.Conv), ('poco_conv_chain', _lib.ConvChain),
             ('poco_basic_block', _lib.BasicBlock), ('poco_bottleneck_tail', _lib.BottleneckTail), ('poco_branch', _lib.Branch),
             ('poco_pack_image', _lib.PackImage), ('poco_fuse_sum', _lib.FuseSum), ('poco_upsample2x', _lib.Upsample2x),
             ('poco_maxpool', _lib.MaxPool), ('poco_avgpool', _lib.AvgPool), ('poco_unpack', _lib.Unpack),
             ('poco_linear', _lib.Linear), ('poco_copy2d', _lib.Copy2d), ('poco_rot6d', _lib.Rot6d),
             ('poco_pare_head', _lib.PareHead), ('poco_realnvp', _lib.RealNVP), ('poco_crop', _lib.Crop),
             ('poco_uncert_post', _lib.UncertPost), ('poco_smpl_model', _lib.SmplModel), ('poco_smpl', _lib.Smpl),
             ('poco_sync', _lib.Sync), ('poco_op', _lib.Op)]
    cname = lambda f: {'in_': 'in'}.get(f, f)      # noqa: E731  (`in` is a Python keyword)
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "poco_b200.h"', 'int main(void) {']
    expect = []
    for c, t in pairs:
        lines.append(f'printf("%zu\\n", sizeof({c}));')
        expect.append((f'sizeof({c})', ctypes.sizeof(t)))
        for f, _ in t._fields_:
            lines.append(f'printf("%zu\\n", offsetof({c}, {cname(f)}));')
            expect.append((f'offsetof({c}, {cname(f)})', getattr(t, f).offset))
    for m, v in (('POCO_MAX_FUSE_INPUTS', _lib.MAX_FUSE_INPUTS), ('POCO_MAX_CHAIN', _lib.MAX_CHAIN),
                 ('POCO_ACT_GUARD_BYTES', _lib.ACT_GUARD_BYTES), ('POCO_SMPL_SCRATCH_FLOATS', _lib.SMPL_SCRATCH_FLOATS),
                 ('POCO_SMPL_DIR_ROWS', _lib.SMPL_DIR_ROWS), ('POCO_OP_SMPL', _lib.OP_SMPL), ('POCO_OP_CONV_CHAIN', _lib.OP_CONV_CHAIN),
                 ('POCO_OP_BASIC_BLOCK', _lib.OP_BASIC_BLOCK), ('POCO_OP_BOTTLENECK_TAIL', _lib.OP_BOTTLENECK_TAIL),
                 ('POCO_OP_BRANCH', _lib.OP_BRANCH), ('POCO_MAX_BRANCH_BLOCKS', _lib.MAX_BRANCH_BLOCKS)):
        lines.append(f'printf("%zu\\n", (size_t){m});')
        expect.append((m, v))
    lines += ['return 0; }']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)],
                   check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert len(got) == len(expect)
    bad = [(name, want, g) for (name, want), g in zip(expect, got) if want != g]
    assert not bad, bad


def test_every_op_rejects_malformed_descriptors_before_touching_the_gpu():
    """error behaviour of the C ABI (include/poco_b200.h): non-zero return + poco_last_error(), never a throw or a
    launch.  Zeroed descriptors for every op kind, then a few op-specific violations."""
    lib = _lib.lib()
    kinds = [_lib.PackImage, _lib.Conv, _lib.ConvChain, _lib.FuseSum, _lib.Upsample2x, _lib.MaxPool, _lib.AvgPool,
             _lib.Unpack, _lib.Linear, _lib.Copy2d, _lib.Rot6d, _lib.PareHead, _lib.RealNVP, _lib.Crop, _lib.UncertPost,
             _lib.Smpl, _lib.BasicBlock, _lib.BottleneckTail, _lib.Branch]
    assert {_lib._KIND_OF_TYPE[t] for t in kinds} == set(_lib._KIND_OF_TYPE.values())
    n0 = _lib.kernel_launches()
    for t in kinds:
        assert lib.poco_run_op(ctypes.byref(_lib.make_op(t())), None) == 1, t.__name__
        assert lib.poco_last_error(), t.__name__
    op = _lib.Op()
    op.kind = 99
    assert lib.poco_run_op(ctypes.byref(op), None) == 1 and b'unknown op kind' in lib.poco_last_error()
    with pytest.raises(_lib.PocoError, match='unknown op kind'):
        _lib.run_op(op, 0)
    plan = ctypes.c_void_p()
    assert lib.poco_plan_create(ctypes.byref(op), 1, ctypes.byref(plan)) == 1
    assert lib.poco_plan_create(None, 0, ctypes.byref(plan)) == 1
    # SMPL stage: the padded vertex count must be a multiple of 128 and the joint table must fit
    buf = ctypes.create_string_buffer(64)
    a = ctypes.addressof(buf)
    m = _lib.SmplModel(a, a, a, a, a, a, a, a, a, a, a, 6890, 6890, 21, 9, 49, 0)
    d = _lib.Smpl(m, a, a, a, 0, 0, 0, 0, 0, 4, 0, 0, 224, 5000.0, 0, a, a, a, a, a, 0)
    assert lib.poco_run_op(ctypes.byref(_lib.make_op(d)), None) == 1 and b'multiple of 128' in lib.poco_last_error()
    d.model.vp = 6912
    d.model.n_extra_vertex = 40
    assert lib.poco_run_op(ctypes.byref(_lib.make_op(d)), None) == 1 and b'too many joints' in lib.poco_last_error()
    d.model.n_extra_vertex = 21
    d.cliff = 1                                 # CLIFF cameras without the box metadata
    assert lib.poco_run_op(ctypes.byref(_lib.make_op(d)), None) == 1 and b'cliff' in lib.poco_last_error()
    # crop: non-positive bbox scale, oversized crop
    c = _lib.Crop(a, 1080, 1920, a, 4, 224, 0.0, 0, a, 0, 0, 0, 0, 0)
    assert lib.poco_run_op(ctypes.byref(_lib.make_op(c)), None) == 1 and b'scale' in lib.poco_last_error()
    c.scale, c.crop = 1.2, 4096
    assert lib.poco_run_op(ctypes.byref(_lib.make_op(c)), None) == 1 and b'geometry' in lib.poco_last_error()
    assert _lib.kernel_launches() == n0, 'a rejected descriptor must not launch anything'


@pytest.mark.parametrize('preset', PRESETS)
def test_state_dict_contract(preset):
    """same names / shapes as the reference (spec_<preset>.json was dumped from pocolib.models.POCO)"""
    meta, _, sd = load_preset(preset)
    m = build_model(preset)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == meta['spec']
    for part in ('backbone', 'head', 'uncert_head', 'flow_head'):
        assert len(list(getattr(m, part).parameters())) > 0        # trainer.py:598-601 per-module groups


@pytest.mark.parametrize('kw,width,view', [
    (dict(loss_ver='norm_flow', sigma_dim=9, nflow_mask_type='alter'), 216, (2, 24, 3, 3)),     # the reference hparams default
    (dict(loss_ver='genG', sigma_dim=9), 48, (2, 48)),
    (dict(loss_ver='gauss', sigma_dim=9), 72, (2, 72)),
    (dict(loss_ver='gauss_sigma', sigma_dim=9), 24, (2, 24)),
])
def test_uncertainty_head_output_width_follows_loss_ver_and_sigma_dim(kw, width, view):
    """poco_head.get_num_uncertainty_outputs (poco_head.py:84-94): the LAST uncert_fc layer is 24 * mult * sigma_dim
    wide, and var_pose is viewed [B, -1, 3, 3] when sigma_dim == 9 (poco_head.py:147-148)"""
    import emu
    from oracle import synth_ckpt as S
    m = POCO(backbone='resnet50-pare', uncert_type=['pose'], smpl_mean_params=S.smpl_mean_params(0), **kw).eval()
    sd = m.state_dict()
    last = sorted(k for k in sd if k.startswith('uncert_head.uncert_fc') and k.endswith('.weight'))[-1]
    assert sd[last].shape[0] == width == m.n_uncert_out
    # the plan writes every column of var_pose (a narrower last layer used to leave zeros behind)
    eng = m._build_engine(2, torch.device('cpu'))
    eng.img.copy_(torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0)))
    emu.run_plan_ops(eng.plan.ops, eng.plan.keep)
    vp = eng.out['var_pose']
    assert vp.shape == (2, width) and (vp != 0).all()
    assert vp.view(2, -1, 3, 3).shape == view if m.var_sigma_dim == 9 else vp.shape == view


def test_submodule_load_state_dict_invalidates_prepared_plans():
    """plans hold folded / packed weight copies: loading into model.backbone must drop them"""
    m = build_model('pare_r50')
    m._engines[(1, 'cpu')] = object()
    m.backbone.load_state_dict(m.backbone.state_dict())
    assert not m._engines
    m._engines[(1, 'cpu')] = object()
    m.head.load_state_dict(m.head.state_dict(), strict=False)
    assert not m._engines


def test_in_place_parameter_updates_invalidate_prepared_plans():
    """model.head.fc1.weight.data.mul_(2) (or any optimiser step) must not leave forward() on stale packed weights"""
    m = build_model('cliff_w32')
    m._build_engine = lambda B, dev: ('plan', B)
    m._engine(4, torch.device('cpu'))
    assert len(m._engines) == 1
    m._engine(4, torch.device('cpu'))
    assert len(m._engines) == 1                         # unchanged parameters: the plan is reused
    with torch.no_grad():
        m.head.deccam.bias.add_(1.0)
    built = []
    m._build_engine = lambda B, dev: built.append(B) or ('plan2', B)
    assert m._engine(4, torch.device('cpu')) == ('plan2', 4) and built == [4]


def test_plan_cache_is_bucketed_and_bounded():
    """tester.py:213 calls forward with B = #detections: batch sizes round up to buckets, plans are LRU-evicted"""
    assert [POCO.bucket(b) for b in (1, 5, 8, 9, 16, 17, 64, 65, 100, 256, 257, 2048)] == \
        [1, 5, 8, 16, 16, 24, 64, 96, 128, 256, 320, 2048]
    m = build_model('pare_r50')
    m.MAX_PLANS = 2
    built = []
    m._build_engine = lambda B, dev: built.append(B) or ('plan', B)
    for B in (4, 16, 4, 24, 16):
        m._engine(B, torch.device('cpu'))
    assert built == [4, 16, 24, 16] and [k[0] for k in m._engines] == [24, 16]


def test_load_pretrained_variants(tmp_path):
    meta, _, sd = load_preset('pare_r50')
    from oracle import synth_ckpt as S
    for wrap in (lambda d: d, lambda d: {'model': d}, lambda d: {'state_dict': {'model.' + k: v for k, v in d.items()}}):
        f = tmp_path / 'ckpt.pt'
        torch.save(wrap(dict(sd)), f)
        m = POCO(**meta['kwargs'], smpl_mean_params=S.smpl_mean_params(0), pretrained=str(f))
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k]), k


def test_constructor_errors_and_no_cpu_path():
    from oracle import synth_ckpt as S
    with pytest.raises(NameError):
        POCO(backbone='vgg16-pare', smpl_mean_params=S.smpl_mean_params(0))
    m = build_model('pare_r50')
    with pytest.raises(PocoError):
        m({'img': torch.zeros(1, 3, 224, 224)})            # CPU tensors: there is no fallback
    with pytest.raises(KeyError):
        m({})
    with pytest.raises(PocoError, match='is_train'):       # training branches (poco_head.py:102, nf_head.py:85) are not built
        m({'img': torch.zeros(1, 3, 224, 224), 'is_train': True})


def test_fold_and_pack_weights():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(32, 16, 3, 3, generator=g)
    bn = (torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g), torch.randn(32, generator=g),
          torch.rand(32, generator=g) + 0.5)
    x = torch.randn(2, 16, 9, 9, generator=g)
    wf, bf = engine.fold_bn(w, None, bn)
    ref = F.batch_norm(F.conv2d(x, w, padding=1), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), ref, atol=1e-4)
    p = engine.pack_conv_weight(wf)
    assert p.shape == (9, 2, 32, 8) and p.dtype == torch.float16
    t, ci, co = 5, 11, 7
    assert p[t, ci // 8, co, ci % 8] == wf[co, ci, t // 3, t % 3].half()


def test_planar_layout_round_trip():
    x = torch.randn(3, 24, 5, 7).half().float()
    a = engine.to_planar(x)
    assert torch.equal(engine.from_planar(a), x)
    v = engine.act_view(a)
    assert v[:, :, 0].abs().sum() == 0 and v[:, :, :, 0].abs().sum() == 0      # zero halo
    assert v[:, :, -1].abs().sum() == 0 and v[:, :, :, -1].abs().sum() == 0


# fp16 storage of weights / activations through ~110 conv layers: what the schedule can deliver
# against the fp32 reference on this synthetic checkpoint (see DESIGN.md "precision").
EMU_TOL = {'pred_pose': 6e-2, 'pred_shape': 6e-3, 'pred_cam': 1.5e-2, 'var_pose': 1e-3}


@pytest.mark.parametrize('preset', ['pare_r50', 'cliff_w32'])
def test_schedule_replay_matches_reference(preset):
    """build the op schedule on CPU buffers, replay it with the op interpreter, compare with the
    reference goldens: validates graph wiring, buffer reuse, BN folding and weight packing."""
    meta, gold, _ = load_preset(preset)
    m = build_model(preset)
    batch = synthetic_batch(preset)
    eng = m._build_engine(meta['test_b'], torch.device('cpu'))
    eng.img.copy_(batch['img'])
    if eng.bbox is not None:
        eng.bbox.copy_(batch['bbox_info'])
    emu.run_plan_ops(eng.plan.ops, eng.plan.keep)
    for k in GATED:
        assert rel_err(eng.out[k].reshape(gold[k].shape).numpy(), gold[k]) < EMU_TOL[k], k
    assert eng.plan.num_ops == len(eng.plan.ops) > 50
    assert eng.plan.flops > 0


def test_fused_kernel_support_predicates():
    """poco_basic_block_supported / poco_bottleneck_tail_supported / poco_branch_supported are pure host functions: which
    geometries the fused kernels take (include/poco_b200.h); everything else falls back to separate poco_conv launches"""
    lib = _lib.lib()
    assert lib.poco_basic_block_supported(32, 56, 56) == 1 and lib.poco_basic_block_supported(32, 56, 61) == 1
    assert lib.poco_basic_block_supported(32, 56, 62) == 0 and lib.poco_basic_block_supported(48, 56, 56) == 0
    assert lib.poco_basic_block_supported(64, 28, 28) == 1 and lib.poco_basic_block_supported(64, 56, 56) == 1
    assert lib.poco_basic_block_supported(64, 28, 62) == 0 and lib.poco_basic_block_supported(128, 14, 14) == 0
    assert lib.poco_bottleneck_tail_supported(64, 256, 56, 56) == 1 and lib.poco_bottleneck_tail_supported(128, 512, 28, 28) == 0
    assert lib.poco_branch_supported(128, 14, 14, 4) == 1 and lib.poco_branch_supported(128, 7, 7, 1) == 1
    assert lib.poco_branch_supported(128, 6, 20, 3) == 1 and lib.poco_branch_supported(128, 15, 15, 4) == 0     # 17 * 17 > 256
    assert lib.poco_branch_supported(64, 14, 14, 4) == 0 and lib.poco_branch_supported(256, 7, 7, 4) == 0
    assert lib.poco_branch_supported(128, 14, 14, 5) == 0 and lib.poco_branch_supported(128, 14, 14, 0) == 0


def test_plan_fuses_blocks_in_fp16_mode_only(monkeypatch):
    """the HRNet-W32 plan: 32 + 29 BasicBlocks (all 32-channel blocks, seven of them also writing the phase-split copy; the
    64-channel blocks except the three that write such a copy), the four Bottleneck tails of layer1 and the seven
    128-channel branches run as fused launches in fp16 mode;
    the parity mode and the POCO_B200_FUSE_* = 0 switches use separate convs; either way every conv of the network is
    accounted for in conv_log"""
    def kinds(m):
        eng = m._build_engine(2, torch.device('cpu'))
        c = {}
        for op in eng.plan.ops:
            c[op.kind] = c.get(op.kind, 0) + 1
        return c, len(eng.plan.conv_log)
    m = build_model('cliff_w32')
    c, n_convs = kinds(m)
    assert (c.get(_lib.OP_BASIC_BLOCK), c.get(_lib.OP_BOTTLENECK_TAIL), c.get(_lib.OP_BRANCH)) == (61, 4, 7)
    assert c[_lib.OP_CONV] + 2 * 61 + 2 * 4 + 8 * 7 == n_convs
    for var in ('POCO_B200_FUSE_BLOCK', 'POCO_B200_FUSE_TAIL', 'POCO_B200_FUSE_BRANCH'):
        monkeypatch.setenv(var, '0')
    c0, n0 = kinds(build_model('cliff_w32'))
    assert n0 == n_convs == c0[_lib.OP_CONV] and not any(k in c0 for k in (_lib.OP_BASIC_BLOCK, _lib.OP_BOTTLENECK_TAIL, _lib.OP_BRANCH))
    for var in ('POCO_B200_FUSE_BLOCK', 'POCO_B200_FUSE_TAIL', 'POCO_B200_FUSE_BRANCH'):
        monkeypatch.delenv(var)
    monkeypatch.setenv('POCO_B200_FUSE_BLOCK64', '0')
    c64, _ = kinds(build_model('cliff_w32'))
    assert c64.get(_lib.OP_BASIC_BLOCK) == 32 and c64.get(_lib.OP_BRANCH) == 7
    monkeypatch.delenv('POCO_B200_FUSE_BLOCK64')
    monkeypatch.setenv('POCO_B200_FUSE_BLOCK_S2D', '0')
    assert kinds(build_model('cliff_w32'))[0].get(_lib.OP_BASIC_BLOCK) == 54
    monkeypatch.delenv('POCO_B200_FUSE_BLOCK_S2D')
    cs, ns = kinds(build_model('cliff_w32', precision='split'))
    assert ns == n_convs == cs[_lib.OP_CONV] and _lib.OP_BRANCH not in cs and _lib.OP_BASIC_BLOCK not in cs


@pytest.mark.parametrize('env', [{'POCO_B200_MERGE_PHASES': '1'}, {'POCO_B200_LANE_ORDER': 'rev', 'POCO_B200_LANE_ORDER_FUSE': '1'},
                                 {'POCO_B200_OUT_LANES': '1'}, {'POCO_B200_LANES': '0'}])
def test_plan_level_switches_keep_the_schedule_correct(env, monkeypatch):
    """the opt-in plan structures (fuse + next branches in one fork / join region, reversed lane emission, output-stage
    lanes, no lanes at all) emit the same network: replayed by the CPU interpreter they reproduce the default schedule's
    outputs exactly, with as many forks as joins and every lane index inside its fork"""
    def run():
        m = build_model('cliff_w32')
        batch = synthetic_batch('cliff_w32')
        eng = m._build_engine(2, torch.device('cpu'))
        eng.img.copy_(batch['img'][:2])
        if eng.bbox is not None:
            eng.bbox.copy_(batch['bbox_info'][:2])
        emu.run_plan_ops(eng.plan.ops, eng.plan.keep)
        return {k: eng.out[k].clone() for k in GATED}, eng.plan.ops
    base, _ = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got, ops = run()
    for k in GATED:
        assert torch.equal(got[k], base[k]), k
    open_lanes = 0
    for op in ops:
        if op.kind == _lib.OP_FORK:
            assert open_lanes == 0
            open_lanes = op.u.sync.n_lanes
        elif op.kind == _lib.OP_JOIN:
            assert open_lanes == op.u.sync.n_lanes
            open_lanes = 0
        else:
            assert op.lane == 0 if open_lanes == 0 else 0 <= op.lane < open_lanes
    assert open_lanes == 0


def test_buffer_pool_keeps_zero_halo():
    """pool reuse never hands a buffer to a tensor of different spatial geometry"""
    b = engine.PlanBuilder({}, 2, 'cpu')
    a = b.act(32, 8, 8)
    b.free(a)
    c = b.act(16, 8, 8)
    assert c.buf is a.buf and c.C == 16
    d = b.act(16, 4, 4)
    assert d.buf is not a.buf
    # concurrent lanes: a buffer freed inside lane 1 is not handed to lane 2 before the join
    b.free(c)
    b.fork([1, 1, 1])
    b.set_lane(1)
    e = b.act(16, 8, 8)
    assert e.buf is a.buf                   # free before the fork: any lane may take it
    b.free(e)
    b.set_lane(2)
    f = b.act(16, 8, 8)
    assert f.buf is not a.buf
    b.join()
    g = b.act(16, 8, 8)
    assert g.buf is a.buf                   # after the join everything is reusable again
    assert sum(b.ops[i].kind in (13, 14) for i in range(len(b.ops))) == 2


def _chain_builder(chained, groups=1, N=4, ch=128, H=14):
    """the plan builder's view of one HRNet branch (4 BasicBlocks) on CPU buffers"""
    from poco_b200 import arch
    g = torch.Generator().manual_seed(3)
    sd = {}
    for k in range(4):
        for c, bn in (('conv1', 'bn1'), ('conv2', 'bn2')):
            sd[f'br.{k}.{c}.weight'] = torch.randn(ch, ch, 3, 3, generator=g) * 0.05
            sd[f'br.{k}.{bn}.weight'] = torch.ones(ch)
            sd[f'br.{k}.{bn}.bias'] = torch.zeros(ch)
            sd[f'br.{k}.{bn}.running_mean'] = torch.zeros(ch)
            sd[f'br.{k}.{bn}.running_var'] = torch.ones(ch)
    b = engine.PlanBuilder(sd, N, 'cpu')
    b.use_chains = chained
    x = engine.alloc_act(ch, N, H, H, 'cpu')
    b.begin_chain()
    for k in range(4):
        x = arch.basic_block(b, x, f'br.{k}', ch, ch)
    b.end_chain(groups=groups)
    return b


def test_conv_chain_descriptors_and_validation():
    """host side of poco_conv_chain: the builder links eight same-geometry convs (segment i reads what i-1
    wrote, residuals come from two segments back or from outside), crop ranges split the batch, and the C ABI
    rejects malformed chains before touching the GPU"""
    b = _chain_builder(True)
    assert [op.kind for op in b.ops] == [_lib.OP_CONV_CHAIN]
    ch = b.ops[0].u.conv_chain
    assert ch.n_seg == 8 and ch.flags
    for i in range(1, 8):
        assert ch.seg[i].in_.data == ch.seg[i - 1].out.data
        assert not ch.seg[i].residual or ch.seg[i].residual != ch.seg[i - 1].out.data
    assert [bool(ch.seg[i].residual) for i in range(8)] == [False, True] * 4
    tiles = (4 * 16 * 16 + 127) // 128
    assert _lib.lib().poco_conv_chain_flag_count(ctypes.byref(ch)) == 7 * tiles
    # crop ranges: 4 crops in 3 groups -> chains over 1, 1 and 2 crops with offset pointers
    b3 = _chain_builder(True, groups=3)
    assert [op.kind for op in b3.ops] == [_lib.OP_CONV_CHAIN] * 3
    ns = [op.u.conv_chain.seg[0].in_.N for op in b3.ops]
    assert ns == [1, 1, 2]
    base = b3.ops[0].u.conv_chain.seg[0].in_.data
    assert [op.u.conv_chain.seg[0].in_.data - base for op in b3.ops] == [0, 16 * 16 * 16, 2 * 16 * 16 * 16]
    # without chains the same branch is eight plain conv ops
    assert [op.kind for op in _chain_builder(False).ops] == [_lib.OP_CONV] * 8
    # malformed chains fail in validation (no GPU needed): bad length, debug-kernel segment
    bad = _lib.ConvChain.from_buffer_copy(ch)
    bad.n_seg = 9
    assert _lib.lib().poco_run_op(ctypes.byref(_lib.make_op(bad)), None) != 0
    assert b'chain length' in _lib.lib().poco_last_error()
    bad = _lib.ConvChain.from_buffer_copy(ch)
    bad.seg[3].impl = 1
    assert _lib.lib().poco_run_op(ctypes.byref(_lib.make_op(bad)), None) != 0
    assert b'tcgen05 kernel only' in _lib.lib().poco_last_error()


def test_weight_layouts():
    """pack_conv_weight / pack_conv_weight_dxn index maps (include/poco_b200.h poco_conv.weight, wfmt)"""
    g = torch.Generator().manual_seed(4)
    w = torch.randn(32, 16, 3, 3, generator=g)
    p = engine.pack_conv_weight(w)                      # [tap][Cin/8][Cout][8]
    assert tuple(p.shape) == (9, 2, 32, 8)
    assert p[5, 1, 7, 3] == w[7, 11, 1, 2].half()       # tap 5 = (r 1, s 2), channel 8 + 3
    q = engine.pack_conv_weight_dxn(w)                  # [r][Cin/8][s*Cout + co][8]
    assert tuple(q.shape) == (3, 2, 96, 8)
    assert q[1, 1, 2 * 32 + 7, 3] == w[7, 11, 1, 2].half()
    assert engine.dxn_applies(32, 32, 3, 1, 1) == (os.environ.get('POCO_B200_DXN', '0') == '1')
    assert not engine.dxn_applies(32, 48, 3, 1, 1)


def test_convert_crop_cam_matches_reference_formula():
    """stream.convert_crop_cam_to_orig_img == the numpy formula of pocolib/utils/demo_utils.py:249-266"""
    from poco_b200 import convert_crop_cam_to_orig_img
    g = torch.Generator().manual_seed(9)
    cam = torch.rand(7, 3, generator=g) + 0.5
    bbox = torch.rand(7, 4, generator=g) * 300 + 50
    W, H = 1920, 1080
    got = convert_crop_cam_to_orig_img(cam, bbox, W, H).numpy()
    c, b = cam.numpy(), bbox.numpy()
    sx = c[:, 0] * (1. / (W / b[:, 2]))
    sy = c[:, 0] * (1. / (H / b[:, 2]))
    ref = np.stack([sx, sy, ((b[:, 0] - W / 2.) / (W / 2.) / sx) + c[:, 1], ((b[:, 1] - H / 2.) / (H / 2.) / sy) + c[:, 2]]).T
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-6) and got.shape == (7, 4)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) prints one JSON line with the keys the
    bench contract names; runs the oracle port (or the reference tree when present) on a 1-crop sample"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--cpu-sample', '1'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'crops/sec' and line['unit'] == 'crops/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and 'model' not in line['config']
