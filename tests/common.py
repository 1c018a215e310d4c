"""shared helpers for the test-suite (test infrastructure; may import oracle/)"""
import functools
import json
import os

import numpy as np
import torch

from oracle import synth_ckpt as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PRESETS = ['pare_r50', 'pare_w32', 'cliff_w32', 'cliff_w48cls']
GATED = ('pred_pose', 'pred_shape', 'pred_cam', 'var_pose')      # the north-star parity outputs


@functools.lru_cache(maxsize=None)
def load_preset(preset):
    meta = json.load(open(os.path.join(GOLD, f'spec_{preset}.json')))
    gold = dict(np.load(os.path.join(GOLD, f'golden_{preset}.npz')))
    calib = np.load(os.path.join(GOLD, f'calib_{preset}.npz'))
    sd = S.synth_state_dict(S.template_from_spec(meta, meta['ckpt_seed']), meta['ckpt_seed'], calib)
    return meta, gold, sd


def build_model(preset, device='cpu', **kw):
    from poco_b200 import POCO
    meta, gold, sd = load_preset(preset)
    m = POCO(**meta['kwargs'], smpl_mean_params=S.smpl_mean_params(meta['ckpt_seed']), **kw)
    m.load_state_dict(sd)
    return m.to(device).eval()


def rel_err(a, ref):
    """max-abs-err / max-abs-ref per tensor"""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def synthetic_batch(preset, device='cpu', B=None):
    meta, _, _ = load_preset(preset)
    return S.synthetic_batch(B or meta['test_b'], meta['test_seed'], device)
