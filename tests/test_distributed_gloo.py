"""N>1 host logic on CPU (gloo, world_size 2): shard partition, unique-id style broadcast, and the
identity "all-reduce of per-shard sums == sums over the whole source" that the sharded ICP relies on
(checked with the oracle's estimator on each rank's block).  No GPU, no product compute."""
import os
import socket

import numpy as np
import pytest

from cupoch_b200.distributed import gather_correspondences, shard_range

# torch is imported inside the tests: collecting this module for `-m gpu` must not pay the cold `import torch`


def test_shard_range_tiles_exactly():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_py as orc
    from cupoch_b200.testing import datagen
    n = 6000
    tgt, tn = datagen.surface(n, 11)
    src = datagen.make_source(tgt, datagen.gt_transform((-0.5, 0.7, 1.0), (0.004, -0.002, 0.003)), 13, 14, 3e-4)
    # rank 0 "creates the id", everybody must end up with the same bytes (stand-in for the NCCL unique id)
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(buf, 0)
    assert buf.tolist() == list(range(128))
    lo, hi = shard_range(n, rank, world)
    corr, fit, rmse = orc.correspondences(src[lo:hi], tgt, 0.03)
    sums = orc.jtj_jtr(orc.P2PLANE, src[lo:hi], tgt, corr, tgt_nrm=tn)
    t = torch.from_numpy(np.concatenate([sums[:28], [float(len(corr))]]))
    dist.all_reduce(t)                       # the one collective of the sharded loop
    full_corr = gather_correspondences(dist, corr, lo, world)
    if rank == 0:
        q.put((t.numpy(), full_corr))
    dist.destroy_process_group()


def test_sharded_sums_equal_global_sums(orc):
    import torch.multiprocessing as mp
    from cupoch_b200.testing import datagen
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    red, full_corr = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    n = 6000
    tgt, tn = datagen.surface(n, 11)
    src = datagen.make_source(tgt, datagen.gt_transform((-0.5, 0.7, 1.0), (0.004, -0.002, 0.003)), 13, 14, 3e-4)
    corr, _, _ = orc.correspondences(src, tgt, 0.03)
    sums = orc.jtj_jtr(orc.P2PLANE, src, tgt, corr, tgt_nrm=tn)
    np.testing.assert_array_equal(full_corr, corr)             # rank-order concatenation == single-GPU list
    assert red[28] == len(corr)
    np.testing.assert_allclose(red[:28], sums[:28], rtol=1e-12, atol=1e-14 * np.abs(sums[:28]).max())
    # float64 sums agree to ~1e-16: the float32 system every rank solves is identical
    a, b = red[:27].astype(np.float32), sums[:27].astype(np.float32)
    assert (np.abs(a.view(np.int32) - b.view(np.int32)) <= 1).all()


def _normals_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_py as orc
    from cupoch_b200 import distributed
    from cupoch_b200.testing import datagen
    pts, _ = datagen.surface(5001, 17)                      # odd size: unequal blocks

    class Cloud:                                            # what the orchestration needs of a PointCloud
        def __len__(self):
            return len(pts)

    def local(cloud, first, count):                         # stand-in for cphb_estimate_normals_range
        idx, _, _ = orc.search(pts, pts[first:first + count], 20, kdtree=True)
        full = np.full((len(pts), 20), -1, np.int32)
        full[:count] = idx                                  # rows index the whole cloud
        return orc.normals_from_neighbors(pts, full)[:count]

    c = Cloud()
    full = distributed.estimate_normals(c, None, dist, rank, world, local_fn=local)
    if rank == 1:
        q.put(full)
    dist.destroy_process_group()


def test_sharded_estimate_normals_equals_single(orc):
    """replicate the index, shard the queries, all-gather the per-point outputs -- the gathered normals
    are the single-process normals bit for bit (world 2, blocks of unequal size)."""
    import torch.multiprocessing as mp
    from cupoch_b200.testing import datagen
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_normals_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pts, _ = datagen.surface(5001, 17)
    ref = orc.estimate_normals(pts, knn=20)
    np.testing.assert_array_equal(got, ref)


class _CpuVoxelOps:
    """CPU stand-in for the three library calls of distributed.voxel_down_sample (the oracle's arithmetic)"""

    def __init__(self, orc):
        self.orc = orc

    def bounds(self, points):
        p = points.numpy()
        return self.orc.min_bound(p), self.orc.max_bound(p)

    def indices(self, points, voxel, origin):
        import torch
        p = points.numpy()
        return torch.from_numpy(np.floor((p - np.asarray(origin, np.float32)) / np.float32(voxel)).astype(np.int32))

    def down_sample(self, points, normals, colors, voxel, origin):
        import torch
        op, on, oc = self.orc.voxel_down_sample(points.numpy(), float(voxel), None if normals is None else normals.numpy(),
                                                None if colors is None else colors.numpy(), origin=origin)
        f = (lambda a: None if a is None else torch.from_numpy(a))
        return f(op), f(on), f(oc)


def _voxel_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_py as orc
    from cupoch_b200 import distributed
    rng = np.random.default_rng(5)
    pts = (rng.random((20011, 3), dtype=np.float32) * np.array([4, 4, 1], np.float32)).astype(np.float32)
    nrm = rng.standard_normal((20011, 3)).astype(np.float32)
    col = rng.random((20011, 3), dtype=np.float32)
    mine = np.arange(len(pts)) % world == rank            # an interleaved (non-spatial) partition of the input
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[mine]))
    out = distributed.voxel_down_sample(t(pts), 0.11, dist, rank, world, normals=t(nrm), colors=t(col), ops=_CpuVoxelOps(orc))
    slab = distributed.voxel_down_sample(t(pts), 0.11, dist, rank, world, ops=_CpuVoxelOps(orc), gather=False)
    none = distributed.voxel_down_sample(t(pts), 0.0, dist, rank, world, ops=_CpuVoxelOps(orc))
    whole = lambda a: torch.from_numpy(a)
    rep = distributed.voxel_down_sample(whole(pts), 0.11, dist, rank, world, normals=whole(nrm), colors=whole(col),
                                        ops=_CpuVoxelOps(orc), replicated=True)
    for a, b in zip(rep, out):
        assert torch.equal(a, b)                          # the replicated-input path gives the same cloud
    q.put((rank, [o.numpy() for o in out], slab[0].numpy(), none[0].shape[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_voxel_down_sample_equals_single(orc, world):
    """common grid by all-reduce, slabs of x-indices, one all-to-all of the points, the ordinary kernel per
    slab; rank-order concatenation == the single-process VoxelDownSample, bit for bit (points, normals, colours)."""
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_voxel_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get() for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    pts = (rng.random((20011, 3), dtype=np.float32) * np.array([4, 4, 1], np.float32)).astype(np.float32)
    nrm = rng.standard_normal((20011, 3)).astype(np.float32)
    col = rng.random((20011, 3), dtype=np.float32)
    rp, rn, rc = orc.voxel_down_sample(pts, 0.11, nrm, col)
    for rank, out, slab, n_none in res:
        np.testing.assert_array_equal(out[0], rp)
        np.testing.assert_array_equal(out[1], rn)
        np.testing.assert_array_equal(out[2], rc)
        assert n_none == 0
    np.testing.assert_array_equal(np.concatenate([r[2] for r in res]), rp)      # slabs in rank order
    sizes = [len(r[2]) for r in res]
    assert min(sizes) > 0.5 * max(sizes)                                          # slabs are balanced
