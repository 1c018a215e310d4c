"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (saidwivedi/POCO) on CPU.

The reference is pure Python/PyTorch but cannot be imported as shipped in this image (missing
yacs / flatten_dict / smplx / pytorch_lightning, removed torchvision.models.utils, licence-gated
data files).  This module injects tiny import shims into ``sys.modules`` (no reference code is
copied), synthesises ``data/smpl_mean_params.npz`` in a temp cwd, stubs the SMPL mesh stage and
then builds ``pocolib.models.POCO`` exactly the way ``pocolib/core/tester.py:75-98`` does.

It is used by
  * ``oracle/make_golden.py``  -- generates tests/golden/*.npz from the reference forward,
  * ``bench.py --impl reference`` / the ``cpu_baseline`` leg when the reference tree is present
    (``$POCO_REF`` -> ``baseline/_ref`` -> ``/root/reference``).
Nothing under ``poco_b200/`` imports it.
"""
import contextlib
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn
import yaml

from .synth_ckpt import smpl_mean_params, synthetic_batch  # noqa: F401  (re-exported)

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def find_reference_root():
    """$POCO_REF -> baseline/_ref -> /root/reference (first that holds pocolib/)."""
    cands = [os.environ.get('POCO_REF'), os.path.join(_REPO, 'baseline', '_ref'), '/root/reference']
    for c in cands:
        if c and os.path.isdir(os.path.join(c, 'pocolib')):
            return c
    return None


# ----------------------------------------------------------------------------------------------
# shims
# ----------------------------------------------------------------------------------------------
class _CfgNode(dict):
    """Just enough of yacs.config.CfgNode for pocolib/core/config.py and the HRNet cfg builders
    (attribute access, clone, merge_from_file)."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        out = _CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, _CfgNode) else (list(v) if isinstance(v, list) else v)
        return out

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), _CfgNode):
                self[k]._merge(v)
            else:
                self[k] = _CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_other_cfg(self, other):
        self._merge(other)


def _install_shims():
    if 'yacs.config' not in sys.modules:
        yacs = types.ModuleType('yacs')
        yacs_config = types.ModuleType('yacs.config')
        yacs_config.CfgNode = _CfgNode
        yacs.config = yacs_config
        sys.modules['yacs'] = yacs
        sys.modules['yacs.config'] = yacs_config
    if 'flatten_dict' not in sys.modules:
        fd = types.ModuleType('flatten_dict')
        fd.flatten = lambda d, **k: d
        fd.unflatten = lambda d, **k: d
        sys.modules['flatten_dict'] = fd
    if 'pytorch_lightning' not in sys.modules:
        pl = types.ModuleType('pytorch_lightning')
        sys.modules['pytorch_lightning'] = pl
    if 'smplx' not in sys.modules:
        smplx = types.ModuleType('smplx')
        body_models = types.ModuleType('smplx.body_models')
        lbs = types.ModuleType('smplx.lbs')

        class SMPL(nn.Module):  # import-only placeholder; the mesh stage is stubbed below
            def __init__(self, *a, **k):
                super().__init__()

        class SMPLOutput(dict):
            pass

        smplx.SMPL = SMPL
        body_models.SMPLOutput = SMPLOutput
        lbs.vertices2joints = lambda J, v: torch.einsum('bik,ji->bjk', v, J)
        smplx.body_models = body_models
        smplx.lbs = lbs
        sys.modules['smplx'] = smplx
        sys.modules['smplx.body_models'] = body_models
        sys.modules['smplx.lbs'] = lbs
    try:
        import torchvision.models.utils  # noqa: F401  (removed upstream; resnet.py:3 needs it)
    except Exception:
        import torchvision.models as tvm
        m = types.ModuleType('torchvision.models.utils')
        m.load_state_dict_from_url = lambda *a, **k: {}
        sys.modules['torchvision.models.utils'] = m
        tvm.utils = m


class StubSMPLStage(nn.Module):
    """Stands in for smpl_head / smplcam_head (smplx + licence-gated model files are absent).
    flow_head only reads .shape[0] / .device of 'smpl_vertices' (nf_head.py:80-81)."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, rotmat, shape, cam, **kw):
        B = rotmat.shape[0]
        z = rotmat.new_zeros
        return {'smpl_vertices': z(B, 6890, 3), 'smpl_joints3d': z(B, 49, 3),
                'smpl_joints2d': z(B, 49, 2), 'pred_cam_t': z(B, 3)}


@contextlib.contextmanager
def _ref_cwd(seed=0):
    """cwd holding data/smpl_mean_params.npz (relative path at config.py:37)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, 'data'))
        np.savez(os.path.join(d, 'data', 'smpl_mean_params.npz'), **smpl_mean_params(seed))
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


_IMPORTED = {}


def import_reference():
    """Returns the reference's pocolib.models.poco module (shimmed), or raises if absent."""
    if 'mod' in _IMPORTED:
        return _IMPORTED['mod']
    root = find_reference_root()
    if root is None:
        raise FileNotFoundError('reference tree not found ($POCO_REF, baseline/_ref, /root/reference)')
    _install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with _ref_cwd():
            import pocolib.models.poco as P
    P.smpl_head = StubSMPLStage
    P.smplcam_head = StubSMPLStage
    _IMPORTED['mod'] = P
    _IMPORTED['root'] = root
    return P


# model presets: (yaml, backbone override, patch cliff 480)
PRESETS = {
    'pare_r50':      ('demo_poco_pare.yaml',  'resnet50-pare',       False),   # BASELINE config 1
    'pare_w32':      ('demo_poco_pare.yaml',  None,                  False),   # BASELINE config 3
    'cliff_w32':     ('demo_poco_cliff.yaml', 'hrnet_w32-cliff',     True),    # BASELINE config 2/4
    'cliff_w48cls':  ('demo_poco_cliff.yaml', None,                  False),   # shipped demo config
}


def preset_kwargs(preset, cfg_dir=None):
    """POCO(**kwargs) for a preset, read from the demo YAML the way tester.py:75-98 does.
    Works without the reference tree when cfg values are given by poco_b200.configs."""
    yaml_name, bb, _ = PRESETS[preset]
    root = find_reference_root()
    P = import_reference()
    from pocolib.core.config import update_hparams
    cfg = update_hparams(os.path.join(root, 'configs', yaml_name))
    c = cfg.POCO
    return dict(
        backbone=bb or c.BACKBONE, img_res=cfg.DATASET.IMG_RES, uncert_layer=c.UNCERT_LAYER,
        activation_type=c.ACTIVATION_TYPE, uncert_type=c.UNCERT_TYPE, uncert_inp_type=c.UNCERT_INP_TYPE,
        loss_ver=c.LOSS_VER, num_neurons=c.NUM_NEURONS, num_flow_layers=c.NUM_FLOW_LAYERS,
        sigma_dim=c.SIGMA_DIM, num_nf_rv=c.NUM_NF_RV, mask_params_id=c.MASK_PARAMS_ID,
        nflow_mask_type=c.NFLOW_MASK_TYPE, exclude_uncert_idx=c.EXCLUDE_UNCERT_IDX,
        use_dropout=c.USE_DROPOUT, use_iter_feats=c.USE_ITER_FEATS, cond_nflow=c.COND_NFLOW,
        context_dim=c.CONTEXT_DIM, gt_pose_cond=c.GT_POSE_COND, gt_pose_cond_ratio=c.GT_POSE_COND_RATIO)


def build_reference(preset, seed=0):
    """Unmodified reference POCO for a preset (eval mode, CPU, reference random init).
    'cliff_w32' applies the documented one-line deviation: the reference hard-codes
    cliff_head.get_output_channels()==2048 (cliff_head.py:129-132), which makes hrnet_w32-cliff
    crash; we return the head's true input width (480) instead."""
    P = import_reference()
    kw = preset_kwargs(preset)
    patch = PRESETS[preset][2]
    import warnings
    from pocolib.models.head import cliff_head as _ch_mod  # noqa: F401
    ch_cls = P.cliff_head
    orig = ch_cls.get_output_channels
    if patch:
        ch_cls.get_output_channels = lambda self: self.num_input_features
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            with _ref_cwd(seed):
                torch.manual_seed(seed)
                model = P.POCO(pretrained=None, **kw)
    finally:
        ch_cls.get_output_channels = orig
    model.eval()
    return model, kw
