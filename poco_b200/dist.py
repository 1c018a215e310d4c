"""Multi-GPU plumbing: crops are independent, so a batch is sharded contiguously across the ranks of
one NVSwitch box (one process per GPU) and the per-crop results are merged with ONE all-gather of a
packed fp32 record -- the only collective on the path.  The reference has no
inference-time communication (tester.py:61 is single-GPU); this is the B200-side scale-out."""
import torch
import torch.distributed as dist

# packed per-crop record: the four parity outputs of the north star
RECORD_FIELDS = (('pred_pose', (24, 3, 3)), ('pred_shape', (10,)), ('pred_cam', (3,)), ('var_pose', (24,)))
RECORD_WIDTH = sum(int(torch.tensor(s).prod()) for _, s in RECORD_FIELDS)      # 253 floats
# optional mesh-stage fields: joints always fit a latency-bound gather (49*5 floats per crop); the
# vertices (6890*3 floats = 83 KB per crop, 21 MB per rank at 256 crops) are a bandwidth-sized second payload
MESH_FIELDS = (('smpl_joints3d', (49, 3)), ('smpl_joints2d', (49, 2)), ('pred_cam_t', (3,)))
MESH_VERTEX_FIELD = (('smpl_vertices', (6890, 3)),)


def record_fields(mesh=False, vertices=False):
    return RECORD_FIELDS + (MESH_FIELDS if mesh or vertices else ()) + (MESH_VERTEX_FIELD if vertices else ())


def shard_range(total, rank, world):
    """contiguous shard [lo, hi) of `total` crops for `rank` (earlier ranks take the remainder)"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_record(out, fields=RECORD_FIELDS):
    """dict of per-crop outputs -> [B, 253] contiguous fp32 (wider with the mesh fields of `record_fields`)"""
    B = out['pred_pose'].shape[0]
    return torch.cat([out[k].reshape(B, -1).float() for k, _ in fields], dim=1).contiguous()


def unpack_record(rec, fields=RECORD_FIELDS):
    out, c = {}, 0
    for k, shape in fields:
        n = 1
        for s in shape:
            n *= s
        out[k] = rec[:, c:c + n].reshape(rec.shape[0], *shape)
        c += n
    return out


def all_gather_outputs(out, group=None, fields=RECORD_FIELDS):
    """one all-gather (NCCL over NVLink on GPUs; gloo in CPU tests) of the packed records.
    Every rank must hold the same number of crops (weak scaling: fixed crops per GPU)."""
    rec = pack_record(out, fields)
    world = dist.get_world_size(group)
    full = torch.empty(world * rec.shape[0], rec.shape[1], dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(full, rec, group=group)
    return unpack_record(full, fields)
