"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/uncert_golden.npz from the reference's own
POCOUtils.prepare_uncert / get_global_uncert (needs the reference tree):   python -m oracle.make_golden_uncert"""
import os
import types

import numpy as np
import torch

from .make_golden_crop import import_reference_utils

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'uncert_golden.npz')


def main():
    import_reference_utils()
    import pocolib.utils.poco_utils as P
    rng = np.random.default_rng(3)
    var = rng.uniform(0.0, 1.0, size=(64, 24)).astype(np.float32)
    var[::7, 0] = 0.95          # rows over both thresholds
    var[3::7, 0] = 0.6          # over the PARE threshold only
    out = {'var': var}
    for bb in ('hrnet_w48_cls-cliff', 'hrnet_w32-pare'):
        for kin in (False, True):
            fake = types.SimpleNamespace(HPS_BACKBONE=bb, LOSS_VER='norm_flow_res_gaus', KINEMATIC_UNCERT=kin)
            prepared = P.POCOUtils.prepare_uncert(fake, torch.from_numpy(var.copy()))        # tester.py:243
            thr = prepared.copy()
            glob = P.POCOUtils.get_global_uncert(fake, thr)                                   # tester.py:244 (in place)
            tag = f"{'cliff' if 'cliff' in bb else 'pare'}_{int(kin)}"
            out[f'prepared_{tag}'] = prepared
            out[f'thresholded_{tag}'] = thr
            out[f'global_{tag}'] = np.asarray(glob, dtype=np.float32)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
