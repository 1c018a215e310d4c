"""SMPL mesh stage: oracle properties + reference-generated camera goldens + host-side model
preparation on the CPU; the CUDA op (poco_smpl_run through the C ABI) against the oracle on the GPU.
LBS parity is UNPINNED (smplx and the licensed model files are absent, oracle/smpl_oracle.py): what can be pinned --
the in-tree camera / projection math -- is, against tests/golden/smpl_cam_golden.npz (oracle/make_golden_smpl.py)."""
import functools
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from oracle import smpl_oracle as O
from poco_b200 import PocoError
from poco_b200 import _lib as L
from poco_b200 import smpl as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'smpl_cam_golden.npz')


@functools.lru_cache(maxsize=None)
def model(seed=0, nv=O.NV):
    return O.synthetic_model(seed, nv)


def random_rotmats(rng, n, scale=1.0):
    """proper rotations from axis-angle vectors ~ N(0, scale) (Rodrigues), [n,24,3,3] float32"""
    aa = rng.standard_normal((n, 24, 3)) * scale
    th = np.linalg.norm(aa, axis=-1, keepdims=True)
    k = aa / np.maximum(th, 1e-12)
    Kx = np.zeros((n, 24, 3, 3))
    Kx[..., 0, 1], Kx[..., 0, 2], Kx[..., 1, 0] = -k[..., 2], k[..., 1], k[..., 2]
    Kx[..., 1, 2], Kx[..., 2, 0], Kx[..., 2, 1] = -k[..., 0], -k[..., 1], k[..., 0]
    th = th[..., None]
    return (np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)).astype(np.float32)


def inputs(n, seed=5, cliff=False):
    r = np.random.default_rng(seed)
    d = {'rotmat': random_rotmats(r, n, 0.4), 'shape': r.standard_normal((n, 10)).astype(np.float32),
         'cam': np.stack([r.uniform(0.5, 1.3, n), r.normal(0, 0.1, n), r.normal(0, 0.1, n)], -1).astype(np.float32)}
    if cliff:
        d.update(img_w=r.choice([640., 1280., 1920.], n).astype(np.float32), img_h=r.choice([480., 720., 1080.], n).astype(np.float32),
                 bbox_center=np.stack([r.uniform(100, 500, n), r.uniform(100, 400, n)], -1).astype(np.float32),
                 bbox_scale=r.uniform(0.6, 3.0, n).astype(np.float32))
        d['focal'] = np.sqrt(d['img_w'] ** 2 + d['img_h'] ** 2).astype(np.float32)
    return d


def oracle_stage(m, d, cliff, normalize=False):
    kw = {k: d[k] for k in ('focal', 'bbox_scale', 'bbox_center', 'img_w', 'img_h')} if cliff else {}
    return O.smpl_stage(m, d['rotmat'], d['shape'], d['cam'], cliff, normalize_joints2d=normalize, **kw)


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_rest_pose_is_the_shaped_template():
    m = model()
    eye = np.broadcast_to(np.eye(3, dtype=np.float32), (2, 24, 3, 3))
    v, j = O.lbs(m, np.zeros((2, 10), np.float32), eye)
    np.testing.assert_allclose(v[0], m['v_template'], atol=1e-12)
    np.testing.assert_allclose(j[0], m['J_regressor'].astype(np.float64) @ m['v_template'].astype(np.float64), atol=1e-12)
    betas = np.random.default_rng(1).standard_normal((2, 10)).astype(np.float32)
    v, j = O.lbs(m, betas, eye)
    shaped = m['v_template'] + np.einsum('bl,mkl->bmk', betas.astype(np.float64), m['shapedirs'].astype(np.float64))
    np.testing.assert_allclose(v, shaped, atol=1e-12)
    np.testing.assert_allclose(j, np.einsum('bik,ji->bjk', shaped, m['J_regressor'].astype(np.float64)), atol=1e-12)


def test_oracle_root_rotation_moves_the_mesh_rigidly():
    """global_orient only enters the root transform [R0 | J0]: replacing R0 by Q R0 maps v -> Q (v - J0) + J0"""
    m = model()
    d = inputs(3)
    v, j = O.lbs(m, d['shape'], d['rotmat'])
    Q = random_rotmats(np.random.default_rng(9), 3, 1.0)[:, 0].astype(np.float64)
    rot2 = d['rotmat'].astype(np.float64).copy()
    rot2[:, 0] = Q @ rot2[:, 0]
    v2, j2 = O.lbs(m, d['shape'], rot2)
    eye = np.broadcast_to(np.eye(3), (3, 24, 3, 3))
    J0 = O.lbs(m, d['shape'], eye)[1][:, 0]                           # rest root joint
    # (exact up to |sum_j w_j - 1| ~ 4e-8 of the float32 skinning weights)
    np.testing.assert_allclose(v2, np.einsum('brc,bvc->bvr', Q, v - J0[:, None]) + J0[:, None], atol=1e-7)
    np.testing.assert_allclose(j2, np.einsum('brc,bvc->bvr', Q, j - J0[:, None]) + J0[:, None], atol=1e-10)


def test_oracle_joints_selection():
    m = model()
    d = inputs(2)
    v, j = O.lbs(m, d['shape'], d['rotmat'])
    J = O.smpl_joints(m, v, j)
    assert J.shape == (2, 49, 3)
    np.testing.assert_array_equal(J[:, 8], j[:, 0])                                        # OP MidHip = SMPL joint 0
    np.testing.assert_array_equal(J[:, 0], v[:, m['extra_vertex_ids'][0]])                 # OP Nose = vertex joint 24
    np.testing.assert_allclose(J[:, 27], np.einsum('bik,i->bk', v, m['J_regressor_extra'][0].astype(np.float64)))  # 45


def test_joint_map_is_the_reference_one():
    ref = np.load(GOLD)['joint_map']
    np.testing.assert_array_equal(ref, O.JOINT_MAP)
    np.testing.assert_array_equal(ref, S.SMPL_JOINT_MAP)


def test_oracle_cameras_match_the_reference_functions():
    g = np.load(GOLD)
    cam, joints = g['cam'], g['joints'].astype(np.float64)
    crop_t = O.weak_perspective_to_perspective(cam)
    np.testing.assert_allclose(crop_t, g['pare_cam_t'], rtol=2e-6)
    np.testing.assert_allclose(O.project(joints, crop_t, 5000., 0., 0.), g['pare_joints2d'], rtol=1e-5, atol=1e-3)
    full_t = O.crop_cam_to_full_img_cam(cam, g['scale'].astype(np.float64) * 200., g['center'].astype(np.float64),
                                        g['img_w'].astype(np.float64), g['img_h'].astype(np.float64), g['focal'].astype(np.float64))
    np.testing.assert_allclose(full_t, g['cliff_full_t'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.project(joints, full_t, g['focal'], g['img_w'] / 2., g['img_h'] / 2.), g['cliff_joints2d'],
                               rtol=1e-5, atol=2e-3)


def test_torch_cameras_match_the_reference_functions():
    """the plain-torch conversions the stub / smplx stages use (poco_b200/smpl.py)"""
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k])    # noqa: E731
    np.testing.assert_allclose(S.weak_perspective_to_perspective(t('cam')).numpy(), g['pare_cam_t'], rtol=1e-6)
    full = S.crop_cam_to_full_img_cam(t('cam'), t('scale') * 200., t('center'), t('img_w'), t('img_h'), t('focal'))
    np.testing.assert_allclose(full.numpy(), g['cliff_full_t'], rtol=1e-6)
    K = S._intrinsics(16, 'cpu', t('focal'), t('img_w') / 2., t('img_h') / 2.)
    np.testing.assert_allclose(S.project(t('joints'), full, K).numpy(), g['cliff_joints2d'], rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------------------------- CPU: host-side preparation
def emulate_kernels(p, rotmat, betas):
    """float64 walk through smpl_pose_kernel / smpl_skin_kernel / smpl_joints_kernel (poco_b200/csrc/smpl.cu) with the
    very arrays and index formulas the device code uses -- validates the layout prepare_smpl_model produces"""
    f = np.float64
    nv, vp = p['nv'], p['vp']
    n = rotmat.shape[0]
    R = rotmat.astype(f).reshape(n, 216)
    verts = np.zeros((n, nv, 3))
    allj = []
    dirs, vt, W = p['dirs'].astype(f).reshape(224 * 3 * vp), p['v_template'].astype(f).reshape(3 * vp), p['weights'].astype(f).reshape(24 * vp)
    for b in range(n):
        coef = np.zeros(220)
        coef[:10] = betas[b]
        for i in range(10, 217):
            e = i - 10
            coef[i] = R[b, 9 + e] - (1.0 if e % 9 in (0, 4, 8) else 0.0)
        J = p['j_template'].astype(f) + p['j_dirs'].astype(f) @ betas[b].astype(f)
        G = np.zeros((24, 12))
        for j in range(24):
            par = p['parents'][j]
            for r in range(3):
                G[j, r * 4:r * 4 + 3] = R[b, j * 9 + r * 3:j * 9 + r * 3 + 3]
                G[j, r * 4 + 3] = J[j, r] - (J[par, r] if par >= 0 else 0.0)
        for i in range(1, 24):
            Gp, Li, out = G[p['parents'][i]], G[i].copy(), np.zeros(12)
            for lane in range(12):
                r, c = lane >> 2, lane & 3
                out[lane] = Gp[r * 4] * Li[c] + Gp[r * 4 + 1] * Li[4 + c] + Gp[r * 4 + 2] * Li[8 + c] + (Gp[r * 4 + 3] if c == 3 else 0.0)
            G[i] = out
        A = G.copy()
        for j in range(24):
            for r in range(3):
                A[j, r * 4 + 3] = G[j, r * 4 + 3] - G[j, r * 4:r * 4 + 3] @ J[j]
        v = np.arange(nv)
        acc = np.stack([vt[c * vp + v] for c in range(3)], -1)
        for k in range(217):
            for c in range(3):
                acc[:, c] += coef[k] * dirs[(k * 3 + c) * vp + v]
        T = np.zeros((nv, 12))
        for j in range(24):
            T += W[j * vp + v][:, None] * A[j][None]
        for r in range(3):
            verts[b, :, r] = T[:, r * 4] * acc[:, 0] + T[:, r * 4 + 1] * acc[:, 1] + T[:, r * 4 + 2] * acc[:, 2] + T[:, r * 4 + 3]
        ja = [G[:, 3::4]]
        ja.append(verts[b, p['extra_vertex_ids']])
        rows = []
        for r in range(len(p['reg_row_ptr']) - 1):
            s = slice(p['reg_row_ptr'][r], p['reg_row_ptr'][r + 1])
            rows.append(p['reg_val'][s].astype(f) @ verts[b, p['reg_col'][s]])
        ja.append(np.array(rows).reshape(-1, 3))
        allj.append(np.concatenate(ja, 0)[p['joint_map']])
    return verts, np.stack(allj)


def test_prepared_layout_reproduces_the_oracle():
    m = model(3, nv=300)                    # (small mesh: the emulation is a python loop)
    p = S.prepare_smpl_model(m)
    assert p['vp'] == 384 and p['dirs'].shape == (224, 3, 384) and not p['dirs'][217:].any() and p['weights'].shape == (24, 384)
    assert not p['dirs'][:, :, 300:].any() and not p['weights'][:, 300:].any()
    d = inputs(2, seed=8)
    v, j = emulate_kernels(p, d['rotmat'], d['shape'])
    ov, oj = O.lbs(m, d['shape'], d['rotmat'])
    np.testing.assert_allclose(v, ov, atol=2e-6)            # (prepared arrays are float32; the fold is done in float64)
    np.testing.assert_allclose(j, O.smpl_joints(m, ov, oj), atol=2e-6)


def test_prepare_accepts_the_pkl_posedirs_layout_and_rejects_bad_models():
    m = dict(model(3, nv=300))
    p0 = S.prepare_smpl_model(m)
    m2 = dict(m, posedirs=m['posedirs'].T.reshape(300, 3, 207))          # official .pkl layout
    np.testing.assert_array_equal(S.prepare_smpl_model(m2)['dirs'], p0['dirs'])
    with pytest.raises(ValueError):
        S.prepare_smpl_model(dict(m, parents=np.array([-1] + [23] * 23)))
    with pytest.raises(ValueError):
        S.prepare_smpl_model(dict(m, joint_map=np.array([0, 99])))
    with pytest.raises(ValueError):
        S.prepare_smpl_model(dict(m, extra_vertex_ids=np.array([300])))
    with pytest.raises(ValueError):
        S.prepare_smpl_model(dict(m, weights=m['weights'][:, :20]))


def test_load_smpl_model_npz_and_chumpy_free_pkl(tmp_path):
    m = model(3, nv=300)
    np.savez(tmp_path / 'm.npz', **m)
    got = S.load_smpl_model(str(tmp_path / 'm.npz'), regressor_extra=None)
    np.testing.assert_array_equal(got['posedirs'], m['posedirs'])
    # an official-style pickle: arrays wrapped in chumpy objects, kintree_table instead of parents
    mod = types.ModuleType('chumpy')
    sub = types.ModuleType('chumpy.ch')

    class Ch:
        def __init__(self, x):
            self.x = x
    Ch.__module__, Ch.__qualname__ = 'chumpy.ch', 'Ch'
    sub.Ch = Ch
    sys.modules['chumpy'], sys.modules['chumpy.ch'] = mod, sub
    try:
        raw = {'v_template': Ch(m['v_template']), 'shapedirs': Ch(m['shapedirs']), 'posedirs': m['posedirs'].T.reshape(300, 3, 207),
               'J_regressor': m['J_regressor'], 'weights': Ch(m['weights']), 'extra_vertex_ids': m['extra_vertex_ids'],
               'kintree_table': np.stack([np.where(m['parents'] < 0, 2 ** 32 - 1, m['parents']).astype(np.int64), np.arange(24)])}
        with open(tmp_path / 'SMPL_NEUTRAL.pkl', 'wb') as fh:
            pickle.dump(raw, fh, protocol=2)
    finally:
        del sys.modules['chumpy'], sys.modules['chumpy.ch']
    np.save(tmp_path / 'extra.npy', m['J_regressor_extra'])
    got = S.load_smpl_model(str(tmp_path / 'SMPL_NEUTRAL.pkl'), regressor_extra=str(tmp_path / 'extra.npy'))
    p = S.prepare_smpl_model(got)
    np.testing.assert_array_equal(p['dirs'], S.prepare_smpl_model(m)['dirs'])
    np.testing.assert_array_equal(p['parents'], m['parents'])


def test_stage_selection_and_no_cpu_path():
    m = model(3, nv=300)
    st = S.make_smpl_stage('cliff', 224, m)
    assert isinstance(st, S.DeviceSmplStage) and st.cliff and st.n_joints_out == 49
    assert isinstance(S.make_smpl_stage('pare', 224, None), (S.StubSmplStage, S.SmplStage))
    assert not [k for k in st.state_dict()], 'model arrays must not leak into the checkpoint namespace'
    with pytest.raises(PocoError):
        st(torch.zeros(1, 24, 3, 3), torch.zeros(1, 10), torch.ones(1, 3))
    assert L.SMPL_SCRATCH_FLOATS == 220 + 24 * 12 + 24 * 3


# ------------------------------------------------------------------------------------------------------ GPU
def run_stage(m, d, cliff, normalize=False):
    st = S.DeviceSmplStage('cliff' if cliff else 'pare', m).to('cuda')
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
    kw = dict(focal_length=t(d['focal']), bbox_scale=t(d['bbox_scale']), bbox_center=t(d['bbox_center']),
              img_w=t(d['img_w']), img_h=t(d['img_h'])) if cliff else dict(normalize_joints2d=normalize)
    n0 = L.kernel_launches()
    out = st(t(d['rotmat']), t(d['shape']), t(d['cam']), **kw)
    torch.cuda.synchronize()
    assert L.kernel_launches() - n0 == 3
    return {k: v.cpu().numpy() for k, v in out.items()}


def check_against_oracle(out, ref):
    assert set(out) == set(ref)
    np.testing.assert_allclose(out['smpl_vertices'], ref['smpl_vertices'], atol=2e-5, rtol=0)
    np.testing.assert_allclose(out['smpl_joints3d'], ref['smpl_joints3d'], atol=2e-5, rtol=0)
    np.testing.assert_allclose(out['pred_cam_t'], ref['pred_cam_t'], rtol=2e-6)
    np.testing.assert_allclose(out['smpl_joints2d'], ref['smpl_joints2d'], rtol=1e-4, atol=2e-2)
    if 'pred_fullimg_cam_t' in ref:
        np.testing.assert_allclose(out['pred_fullimg_cam_t'], ref['pred_fullimg_cam_t'], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('n', [1, 5, 8, 19])
@pytest.mark.parametrize('cliff', [False, True])
def test_gpu_smpl_stage_matches_oracle(n, cliff):
    """ragged batches against the 8-crop CTA tile of the skinning kernel and the 4-crop CTA of the pose kernel"""
    m = model()
    d = inputs(n, seed=20 + n, cliff=cliff)
    check_against_oracle(run_stage(m, d, cliff, normalize=True), oracle_stage(m, d, cliff, normalize=True))


@pytest.mark.gpu
def test_gpu_smpl_small_mesh_and_padding():
    m = model(3, nv=300)                    # vp = 384: the last CTA column is mostly padding
    d = inputs(3, seed=2)
    check_against_oracle(run_stage(m, d, False), oracle_stage(m, d, False))


@pytest.mark.gpu
def test_gpu_smpl_cameras_and_projection_match_reference_golden():
    """the reference's own camera + projection outputs (oracle/make_golden_smpl.py).  The golden joints are fed through
    the kernels with a degenerate model whose LBS is the identity: 49 'vertices' = the joints, one-hot regressors, so
    all three joint sources (LBS joints, vertex joints, extra regressor) are exercised and the arithmetic is exact."""
    g = np.load(GOLD)
    eye = np.eye(49, dtype=np.float32)
    W = np.zeros((49, 24), np.float32)
    W[:, 0] = 1.
    for i in range(16):
        m = {'v_template': g['joints'][i], 'shapedirs': np.zeros((49, 3, 10), np.float32),
             'posedirs': np.zeros((207, 147), np.float32), 'J_regressor': eye[:24], 'weights': W,
             'extra_vertex_ids': np.arange(24, 45), 'J_regressor_extra': eye[45:49], 'joint_map': np.arange(49)}
        d = {'rotmat': np.broadcast_to(np.eye(3, dtype=np.float32), (1, 24, 3, 3)).copy(), 'shape': np.zeros((1, 10), np.float32),
             'cam': g['cam'][i:i + 1]}
        out = run_stage(m, d, False)
        np.testing.assert_array_equal(out['smpl_joints3d'][0, 24:], g['joints'][i, 24:])      # vertex / regressor joints: exact
        np.testing.assert_allclose(out['smpl_joints3d'][0, :24], g['joints'][i, :24], atol=5e-7)  # chain: (J_i - J_p) + G_p
        np.testing.assert_allclose(out['pred_cam_t'][0], g['pare_cam_t'][i], rtol=1e-6)
        np.testing.assert_allclose(out['smpl_joints2d'][0], g['pare_joints2d'][i], rtol=2e-6, atol=1e-3)
        np.testing.assert_allclose(run_stage(m, d, False, normalize=True)['smpl_joints2d'][0], g['pare_joints2d_norm'][i],
                                   rtol=2e-6, atol=1e-5)
        d.update(focal=g['focal'][i:i + 1], bbox_scale=g['scale'][i:i + 1], bbox_center=g['center'][i:i + 1],
                 img_w=g['img_w'][i:i + 1], img_h=g['img_h'][i:i + 1])
        out = run_stage(m, d, True)
        np.testing.assert_allclose(out['pred_fullimg_cam_t'][0], g['cliff_full_t'][i], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(out['pred_cam_t'][0], g['cliff_cam_t'][i], rtol=1e-6)
        np.testing.assert_allclose(out['smpl_joints2d'][0], g['cliff_joints2d'][i], rtol=2e-6, atol=2e-3)


@pytest.mark.gpu
def test_gpu_smpl_properties_at_full_batch():
    """size-independent properties at BASELINE's batch (256 crops): rest pose returns the shaped template, and a root
    rotation moves the mesh rigidly about the rest root joint"""
    m = model()
    n = 256
    d = inputs(n, seed=77)
    rest = dict(d, rotmat=np.broadcast_to(np.eye(3, dtype=np.float32), (n, 24, 3, 3)).copy())
    out = run_stage(m, rest, False)
    shaped = m['v_template'][None] + np.einsum('bl,mkl->bmk', d['shape'], m['shapedirs'])
    np.testing.assert_allclose(out['smpl_vertices'], shaped, atol=2e-6)
    base = run_stage(m, d, False)['smpl_vertices'].astype(np.float64)
    Q = random_rotmats(np.random.default_rng(4), n, 1.0)[:, 0]
    rot2 = d['rotmat'].copy()
    rot2[:, 0] = Q @ rot2[:, 0]
    moved = run_stage(m, dict(d, rotmat=rot2), False)['smpl_vertices']
    J0 = np.einsum('bik,i->bk', shaped.astype(np.float64), m['J_regressor'][0].astype(np.float64))
    np.testing.assert_allclose(moved, np.einsum('brc,bvc->bvr', Q.astype(np.float64), base - J0[:, None]) + J0[:, None], atol=3e-5)


@pytest.mark.gpu
def test_gpu_poco_forward_with_device_smpl_stage():
    """POCO.forward end to end with the mesh stage on the device: keys / shapes of the reference dict, and the mesh
    outputs equal the oracle applied to the forward's own pose / shape / camera"""
    from common import build_model, synthetic_batch
    m = model()
    net = build_model('cliff_w32', 'cuda', smpl_model=m)
    assert isinstance(net.smpl, S.DeviceSmplStage)
    batch = synthetic_batch('cliff_w32', 'cuda')
    with torch.no_grad():
        out = net(batch)
    torch.cuda.synchronize()
    B = batch['img'].shape[0]
    assert out['smpl_vertices'].shape == (B, 6890, 3) and out['smpl_joints3d'].shape == (B, 49, 3)
    assert out['smpl_joints2d'].shape == (B, 49, 2) and out['pred_fullimg_cam_t'].shape == (B, 3)
    c = lambda k: batch[k].cpu().numpy()    # noqa: E731
    d = {'rotmat': out['pred_pose'].cpu().numpy(), 'shape': out['pred_shape'].cpu().numpy(), 'cam': out['pred_cam'].cpu().numpy(),
         'focal': c('focal_length'), 'bbox_scale': c('scale'), 'bbox_center': c('center'), 'img_h': c('orig_shape')[:, 0],
         'img_w': c('orig_shape')[:, 1]}
    ref = oracle_stage(m, d, True)
    got = {k: out[k].cpu().numpy() for k in ref}
    check_against_oracle(got, ref)
