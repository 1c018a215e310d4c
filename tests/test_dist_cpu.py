"""world_size-2 gloo test of the multi-GPU plumbing (contiguous crop shards + ONE all-gather of the
packed per-crop record); runs on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from poco_b200 import dist as pdist


def _fake_outputs(lo, hi):
    idx = torch.arange(lo, hi, dtype=torch.float32)
    return {'pred_pose': idx.view(-1, 1, 1, 1).expand(-1, 24, 3, 3) + 0.25,
            'pred_shape': idx.view(-1, 1).expand(-1, 10) * 2,
            'pred_cam': idx.view(-1, 1).expand(-1, 3) - 1,
            'var_pose': idx.view(-1, 1).expand(-1, 24) / 7}


def _fake_mesh_outputs(lo, hi):
    idx = torch.arange(lo, hi, dtype=torch.float32)
    o = _fake_outputs(lo, hi)
    o.update(smpl_joints3d=idx.view(-1, 1, 1).expand(-1, 49, 3) * 3, smpl_joints2d=idx.view(-1, 1, 1).expand(-1, 49, 2) + 5,
             pred_cam_t=idx.view(-1, 1).expand(-1, 3) * 0.5, smpl_vertices=idx.view(-1, 1, 1).expand(-1, 6890, 3) - 2)
    return o


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = pdist.shard_range(total, rank, world)
    full = pdist.all_gather_outputs(_fake_outputs(lo, hi))
    ref = _fake_outputs(0, total)
    ok = all(torch.equal(full[k], ref[k].contiguous()) for k in ref)
    # the mesh-stage fields ride in the same single collective when asked for
    for vertices in (False, True):
        fields = pdist.record_fields(mesh=True, vertices=vertices)
        full = pdist.all_gather_outputs(_fake_mesh_outputs(lo, hi), fields=fields)
        ref = _fake_mesh_outputs(0, total)
        ok = ok and set(full) == {k for k, _ in fields} and all(torch.equal(full[k], ref[k].contiguous()) for k in full)
    q.put((rank, ok, lo, hi))
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    for total in (1, 7, 8, 2048, 2049):
        for world in (1, 2, 3, 8):
            r = [pdist.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_record_round_trip():
    o = _fake_outputs(0, 5)
    rec = pdist.pack_record(o)
    assert rec.shape == (5, pdist.RECORD_WIDTH) and pdist.RECORD_WIDTH == 253
    back = pdist.unpack_record(rec)
    assert all(torch.equal(back[k], o[k]) for k in o)


def test_mesh_record_round_trip():
    o = _fake_mesh_outputs(0, 3)
    f = pdist.record_fields(mesh=True)
    rec = pdist.pack_record(o, f)
    assert rec.shape == (3, 253 + 49 * 5 + 3)
    back = pdist.unpack_record(rec, f)
    assert 'smpl_vertices' not in back and all(torch.equal(back[k], o[k]) for k in back)
    fv = pdist.record_fields(vertices=True)
    assert pdist.pack_record(o, fv).shape == (3, 253 + 49 * 5 + 3 + 6890 * 3)


def test_all_gather_of_shards_equals_unsharded_gloo_ws2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 8, 8, 16)
