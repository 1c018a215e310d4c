"""Calibrated synthetic checkpoint and synthetic crop batches: DATA ONLY -- no arithmetic of the
hot path lives here.  Used by bench.py, __graft_entry__.smoke(), tools/ and the tests; re-exported as
oracle.synth_ckpt for the golden generators.

The reference's own random init is numerically degenerate for both HRNets (hrnet.py:535 uses
std=0.001 -> features ~1e-10; hrnet_cls.py:504 kaiming fan_out with identity BN -> ~1e8) and the
pretrained checkpoints are licence-gated, so parity is pinned on a *seeded synthetic checkpoint*:

  * every conv / linear / locally-connected weight ~ N(0, gain^2 / fan_in) from a numpy PCG64
    stream seeded by crc32(tensor name) ^ seed  (order independent, platform independent);
  * BatchNorm gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1);
  * BatchNorm running_mean / running_var come from ONE calibration pass of the reference model in
    train mode with momentum=1 on a seeded batch (so every post-BN activation is ~unit variance);
    the resulting statistics are committed as tests/golden/calib_<preset>.npz so the GPU box
    (which has no reference tree) rebuilds the bit-identical checkpoint;
  * structural buffers (flow masks, temperature, init_pose/shape/cam, num_batches_tracked) keep the
    template's values.

The state-dict key names are the reference's (poco.py:131-154 split them by prefix), so the same
file loads into pocolib.models.POCO and into poco_b200.POCO.
"""
import zlib
from collections import OrderedDict

import numpy as np
import torch

_KEEP_SUFFIX = ('.mask', 'temperature', 'init_pose', 'init_shape', 'init_cam', 'num_batches_tracked',
                'mask_params', 'pos_enc')
# decoders the reference initialises with xavier gain 0.01 (cliff_head.py:37-39); give them a real
# gain so pose/shape/cam actually move between crops.
_DECODER_GAIN = {'decpose': 0.3, 'decshape': 0.3, 'deccam': 0.15}


# BatchNorms whose output is ADDED to another path (the closing BN of a residual block, every BN of
# an HRNet fuse layer) get a small gamma, as in trained residual networks: each block is then a small
# perturbation of the identity and the network is not chaotic.  With gamma ~ U(0.5, 1.5) everywhere a
# 1e-7 (fp32 re-association) perturbation of the input grows to 1e-3 at pred_pose, which would make
# any reduced-precision parity test meaningless.
import os as _os
_RES_GAMMA = tuple(float(x) for x in _os.environ.get('POCO_SYNTH_RES_GAMMA', '0.15,0.35').split(','))


def _adds_into_a_sum(bn_prefix, template):
    parent, _, leaf = bn_prefix.rpartition('.')
    if 'fuse_layers' in bn_prefix:
        return True
    if leaf == 'bn3':
        return True
    if leaf == 'bn2' and (parent + '.bn3.weight') not in template and (parent + '.conv2.weight') in template \
            and (parent + '.conv3.weight') not in template and (parent + '.conv1.weight') in template \
            and parent.split('.')[-1].isdigit():
        return True          # BasicBlock (conv1/bn1/conv2/bn2 only)
    return False


def _rng(name, seed):
    return np.random.Generator(np.random.PCG64((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0xFFFFFFFF))


def bn_prefixes(template):
    return [k[:-len('.running_mean')] for k in template.keys() if k.endswith('.running_mean')]


def synth_state_dict(template, seed=0, calib=None):
    """template: OrderedDict name->tensor (a model.state_dict()).  calib: optional dict with
    'names' (list of BN prefixes), 'mean', 'var' (flat float32, concatenated in that order)."""
    bn = set(bn_prefixes(template))
    out = OrderedDict()
    cal = {}
    if calib is not None:
        names = [str(n) for n in calib['names']]
        off = 0
        for n in names:
            c = template[n + '.running_mean'].numel()
            cal[n] = (np.asarray(calib['mean'][off:off + c], np.float32),
                      np.asarray(calib['var'][off:off + c], np.float32))
            off += c
        assert set(names) == bn, 'calibration fixture does not match this model'
    for name, t in template.items():
        if any(name.endswith(s) for s in _KEEP_SUFFIX):
            out[name] = t.clone()
            continue
        prefix, _, leaf = name.rpartition('.')
        r = _rng(name, seed)
        shape = tuple(t.shape)
        if prefix in bn:
            if leaf == 'weight':
                lo, hi = (_RES_GAMMA if _adds_into_a_sum(prefix, template) else (0.5, 1.5))
                v = r.uniform(lo, hi, shape)
            elif leaf == 'bias':
                v = 0.1 * r.standard_normal(shape)
            elif leaf == 'running_mean':
                v = cal[prefix][0] if prefix in cal else np.zeros(shape)
            elif leaf == 'running_var':
                v = cal[prefix][1] if prefix in cal else np.ones(shape)
            else:
                raise KeyError(name)
        elif leaf == 'weight':
            if t.dim() == 4:                       # conv [Cout, Cin, kh, kw]
                fan_in = shape[1] * shape[2] * shape[3]
                std = (2.0 / fan_in) ** 0.5
            elif t.dim() == 2:                     # linear [out, in]
                gain = 1.0
                for key, g in _DECODER_GAIN.items():
                    if prefix.endswith(key):
                        gain = g
                std = gain * (1.0 / shape[1]) ** 0.5
            elif t.dim() == 6:                     # LocallyConnected2d [1, Cout, Cin, J, 1, 1]
                std = (1.0 / shape[2]) ** 0.5
            else:
                raise KeyError(f'unexpected weight rank for {name}: {shape}')
            v = std * r.standard_normal(shape)
        elif leaf == 'bias':
            v = 0.05 * r.standard_normal(shape)
        else:
            raise KeyError(f'do not know how to synthesise {name} {shape}')
        out[name] = torch.from_numpy(np.asarray(v, np.float32).reshape(shape).copy())
    return out


def calibrate_bn(model, batch, run=None):
    """One train-mode pass with momentum=1: running stats := batch stats, layer by layer.
    Returns the calibration dict (names, mean, var)."""
    mods = [(n, m) for n, m in model.named_modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    saved = [(m.momentum, m.training) for _, m in mods]
    was_training = model.training
    model.train()
    for _, m in mods:
        m.momentum = 1.0
    with torch.no_grad():
        (run or model)(batch)
    for (_, m), (mom, _) in zip(mods, saved):
        m.momentum = mom
    model.train(was_training)
    names = [n for n, _ in mods]
    mean = np.concatenate([m.running_mean.detach().cpu().numpy().ravel() for _, m in mods]).astype(np.float32)
    var = np.concatenate([m.running_var.detach().cpu().numpy().ravel() for _, m in mods]).astype(np.float32)
    for _, m in mods:
        m.num_batches_tracked.zero_()
    return {'names': np.array(names), 'mean': mean, 'var': var}


# ------------------------------------------------------------------------------------------------
# helpers that do not need the reference tree (used on the GPU box)
# ------------------------------------------------------------------------------------------------
def smpl_mean_params(seed=0):
    """Synthesised stand-in for the licence-gated data/smpl_mean_params.npz
    (keys pose[144], shape[10], cam[3]; read at cliff_head.py:43-46, pare_head.py:233-236)."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    # identity rotations in the reference's 6-D convention (rot6d reads the vector as a 3x2
    # row-major matrix: a1 = elems 0,2,4), plus small noise so the joints differ
    ident = np.tile(np.array([1, 0, 0, 1, 0, 0], np.float32), 24)
    pose = ident + 0.05 * rng.standard_normal(144).astype(np.float32)
    shape = (0.2 * rng.standard_normal(10)).astype(np.float32)
    cam = np.array([0.9, 0.0, 0.0], np.float32)
    return {'pose': pose.astype(np.float32), 'shape': shape, 'cam': cam}


def synthetic_batch(B, seed=1, device='cpu'):
    """SURVEY 8(d) synthetic inputs (ImageNet-normalised crops are ~N(0,1))."""
    g = torch.Generator().manual_seed(seed)
    batch = {
        'img': torch.randn(B, 3, 224, 224, generator=g),
        'bbox_info': 0.1 * torch.randn(B, 3, generator=g),
        'focal_length': torch.full((B,), 1500.0),
        'scale': torch.ones(B),
        'center': torch.full((B, 2), 500.0),
        'orig_shape': torch.full((B, 2), 1000.0),
    }
    return {k: v.to(device) for k, v in batch.items()}


def alter_masks(num_rv, num_flow_layers):
    """nf_head.py:20-21"""
    a = [i % 2 for i in range(num_rv)]
    b = [(i + 1) % 2 for i in reversed(range(num_rv))]
    return np.array([a, b] * num_flow_layers, np.float32)


def template_from_spec(meta, seed=0):
    """Rebuild a state-dict template (names, shapes, structural buffers) from spec_<preset>.json
    without the reference tree."""
    kw = meta['kwargs']
    mp = smpl_mean_params(seed)
    t = OrderedDict()
    for name, shape in meta['spec'].items():
        leaf = name.rpartition('.')[2]
        if leaf == 'num_batches_tracked':
            v = torch.zeros((), dtype=torch.long)
        elif leaf == 'mask':
            v = torch.from_numpy(alter_masks(kw['num_nf_rv'], kw['num_flow_layers']))
        elif leaf == 'temperature':
            v = torch.tensor(1.0)
        elif leaf == 'init_pose':
            v = torch.from_numpy(mp['pose'][:shape[1]]).unsqueeze(0)
        elif leaf == 'init_shape':
            v = torch.from_numpy(mp['shape']).unsqueeze(0)
        elif leaf == 'init_cam':
            v = torch.from_numpy(mp['cam']).unsqueeze(0)
        else:
            v = torch.zeros(shape)
        assert list(v.shape) == list(shape), (name, v.shape, shape)
        t[name] = v
    return t
