"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the step right after the hot path: the
uncertainty post-processing of the reference tester.  Nothing under poco_b200/ imports it.

Follows pocolib/utils/poco_utils.py: get_kinematic_uncert :21-25 (skeleton kp_utils.py:881-908),
POCOUtils.get_global_uncert :50-61, POCOUtils.prepare_uncert :63-94 (2-D var, the LOSS_VERs of the shipped
configs leave it untouched), as called from pocolib/core/tester.py:243-245 / :418-421.
Parity PINNED: tests/golden/uncert_golden.npz holds the outputs of those reference functions
(oracle/make_golden_uncert.py); this restatement matches them exactly (tests/test_oracle_golden.py)."""
import numpy as np

# (parent, child) rows of get_smpl_skeleton, kp_utils.py:881-908: row i-1 has child i
SMPL_PARENT = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21])


def prepare_uncert(var, kinematic=False, return_conf=False):
    var = np.array(var, dtype=np.float32, copy=True)
    if kinematic:                       # poco_utils.py:21-25: children accumulate their parent's (updated) value
        for i in range(1, 24):
            var[:, i] += var[:, SMPL_PARENT[i]]
    if return_conf:
        var = 1 - var
    return var


def global_uncert(var, backbone, sensitivity_threshold=0.40):
    """-> (var after the in-place thresholding, global value per crop)   poco_utils.py:50-61"""
    var = np.array(var, dtype=np.float32, copy=True)
    if 'cliff' in backbone:
        var[var[:, 0] > 2 * sensitivity_threshold] = 1.0
        return var, var[:, 0].copy()
    var[var[:, 0] > sensitivity_threshold] = 1.0
    return var, var.mean(-1)
