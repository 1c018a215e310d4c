"""Seeded stand-in for the licence-gated SMPL model files (DATA ONLY): arrays of the real dimensions and sparsity for
the mesh-stage tests and benches (data/smpl/SMPL_NEUTRAL.pkl and data/J_regressor_extra.npy cannot be shipped)."""
import numpy as np

NV, NJ, NB = 6890, 24, 10
# SMPL kinematic tree (kintree_table[0] of the published model; the same table poco_utils.py:21-25 walks)
PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int32)
# smplx/vertex_ids.py 'smplh' ids in VertexJointSelector order: face (nose, reye, leye, rear, lear), feet
# (LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel), finger tips left then right (thumb .. pinky)
EXTRA_VERTEX_IDS = np.array([332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                             2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133], dtype=np.int32)
# [constants.JOINT_MAP[n] for n in constants.JOINT_NAMES] (pocolib/core/constants.py:15-93): 49 of the 54
# joints (24 LBS + 21 vertex joints + 9 J_regressor_extra joints)
JOINT_MAP = np.array([24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                      8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27],
                     dtype=np.int32)


def synthetic_model(seed=0, nv=NV, n_extra=9, dtype=np.float32):
    """seeded stand-in for data/smpl/SMPL_NEUTRAL.pkl + data/J_regressor_extra.npy with the real shapes and the
    real sparsity pattern (4 skinning weights per vertex, local joint regressors)"""
    r = np.random.default_rng(seed)
    v_template = (r.standard_normal((nv, 3)) * np.array([0.25, 0.5, 0.12])).astype(dtype)
    shapedirs = (r.standard_normal((nv, 3, NB)) * 0.01).astype(dtype)
    posedirs = (r.standard_normal(((NJ - 1) * 9, nv * 3)) * 0.003).astype(dtype)

    def sparse_rows(rows, k):
        m = np.zeros((rows, nv), dtype=np.float64)
        for i in range(rows):
            idx = r.choice(nv, size=k, replace=False)
            w = r.random(k) + 0.05
            m[i, idx] = w / w.sum()
        return m.astype(dtype)
    J_regressor = sparse_rows(NJ, 40)
    J_regressor_extra = sparse_rows(n_extra, 30)
    weights = np.zeros((nv, NJ), dtype=np.float64)
    for v in range(nv):
        idx = r.choice(NJ, size=4, replace=False)
        w = r.random(4) + 0.05
        weights[v, idx] = w / w.sum()
    return {'v_template': v_template, 'shapedirs': shapedirs, 'posedirs': posedirs, 'J_regressor': J_regressor,
            'weights': weights.astype(dtype), 'parents': PARENTS.copy(),
            'extra_vertex_ids': np.minimum(EXTRA_VERTEX_IDS, nv - 1), 'J_regressor_extra': J_regressor_extra,
            'joint_map': JOINT_MAP.copy()}
