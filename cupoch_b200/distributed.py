"""Multi-GPU plumbing for the sharded ICP: one process per GPU, the source is split
into contiguous blocks, the target is replicated, one all-reduce of 32 doubles per iteration.

torch.distributed is used only for rendezvous (exchanging the CUDA IPC handles or the NCCL unique id);
the communicator itself lives inside libcupoch_b200.so (cphb_comm_*) and is used by the fused loop
(cphb_icp_run(..., comm))."""
import ctypes as C

from . import _lib


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, blocks tile [0, n) in rank order,
    so concatenating per-rank correspondence lists (with lo added to the source index) reproduces the
    single-GPU ordering."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def broadcast_unique_id(dist, rank, device=None):
    """rank 0 creates an NCCL unique id; everybody returns the same 128 bytes (any torch backend)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        b = C.create_string_buffer(128)
        _lib.check(_lib.lib().cphb_nccl_unique_id(b))
        buf = torch.frombuffer(bytearray(b.raw), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def make_comm(dist, rank, world, device=None, kind="p2p"):
    """-> opaque cphb_comm* (ctypes.c_void_p) usable as `comm=` in cupoch_b200.registration.
    kind "p2p": CUDA-IPC mailboxes, the per-iteration exchange is fused into the reduce kernel (NVLink
    stores + flags); kind "nccl": ncclAllReduce."""
    import torch
    L = _lib.lib()
    h = C.c_void_p()
    if kind == "nccl":
        uid = broadcast_unique_id(dist, rank, device)
        _lib.check(L.cphb_comm_nccl_create(uid, world, rank, C.byref(h)))
        return h
    # every rank goes through the same collectives whatever happens locally, then all agree on the outcome
    mine = C.create_string_buffer(64)
    ok = 1
    try:
        _lib.check(L.cphb_comm_p2p_create(world, rank, mine, C.byref(h)))
    except Exception:
        ok = 0
    t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    allh = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    if ok:
        blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
        try:
            _lib.check(L.cphb_comm_p2p_connect(h, blob))
        except Exception:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32)
    if device is not None:
        flag = flag.to(device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if h:
            L.cphb_comm_destroy(h)
        raise _lib.CphbError("peer-memory communicator could not be set up on every rank (CUDA IPC / peer access)")
    return h


def destroy_comm(comm):
    if comm:
        _lib.check(_lib.lib().cphb_comm_destroy(comm))


def gather_correspondences(dist, local_corr, lo, world):
    """all-gather-v of the per-rank (i, j) lists into the single-GPU list, ascending in i (host side, once
    per call).  `lo` is added to the local source indices (0 when the library did the sharding: indices
    are global already)."""
    import numpy as np
    mine = np.asarray(local_corr, np.int32).reshape(-1, 2).copy()
    mine[:, 0] += lo
    out = [None] * world
    dist.all_gather_object(out, mine)
    allc = np.concatenate(out, 0) if out else mine
    return allc[np.argsort(allc[:, 0], kind="stable")]


def all_gather_rows(dist, local_rows, n_total, world, device=None):
    """All-gather of contiguous row blocks (rank r holds rows shard_range(n_total, r, world)) into the full
    [n_total, ...] array, on whatever device the process group works on (gloo: host, nccl: device)."""
    import numpy as np
    import torch
    local_rows = np.ascontiguousarray(local_rows)
    width = local_rows.shape[1:]
    biggest = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    pad = np.zeros((biggest,) + width, local_rows.dtype)
    pad[:len(local_rows)] = local_rows
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.empty((n_total,) + width, local_rows.dtype)
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[lo:hi] = parts[r][:hi - lo].cpu().numpy()
    return out


def _all_gather_row_tensors(dist, local_t, n_total, world):
    """device-side twin of all_gather_rows: torch tensors in, one concatenated tensor out (no host round trip)"""
    import torch
    biggest = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    pad = torch.zeros((max(biggest, 1),) + tuple(local_t.shape[1:]), dtype=local_t.dtype, device=local_t.device)
    pad[:local_t.shape[0]] = local_t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([parts[r][:shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]] for r in range(world)])


def estimate_normals(cloud, search_param, dist, rank, world, device=None, local_fn=None):
    """PointCloud::EstimateNormals over `world` ranks (replicate the index, shard the queries, no
    collective in the search; one all-gather of the per-point outputs).  Every rank holds the full cloud, estimates
    the normals of its contiguous block with cphb_estimate_normals_range and all-gathers the blocks: the result is
    bit-identical to the single-GPU estimate.  With device="cuda" the blocks stay on the device (NCCL all-gather of
    the device buffers, the cloud's normals become a view of the gathered tensor).  `local_fn(cloud, first, count)
    -> [count, 3]` replaces the library call in the gloo/CPU test of this orchestration."""
    import numpy as np
    from . import geometry
    from .utility import DeviceArray
    n = len(cloud)
    lo, hi = shard_range(n, rank, world)
    if local_fn is None:
        sp = search_param or geometry.KDTreeSearchParamKNN()
        if isinstance(sp, geometry.KDTreeSearchParamKNN):
            knn, radius, max_nn = sp.knn, 0.0, 0
        else:
            knn, radius, max_nn = 0, sp.radius, sp.max_nn
        out = DeviceArray((max(hi - lo, 1), 3), np.float32)
        _lib.check(_lib.lib().cphb_estimate_normals_range(cloud._points.ptr, n, knn, radius, max_nn, lo, hi - lo, out.ptr, None))
        if device is not None and str(device).startswith("cuda"):
            import torch
            _lib.check(_lib.lib().cphb_stream_synchronize(None))
            mine_t = torch.as_tensor(out, device="cuda")[:hi - lo]
            full_t = _all_gather_row_tensors(dist, mine_t, n, world)
            cloud.normals = full_t                       # zero-copy: the cloud borrows the gathered tensor
            return full_t
        mine = out.cpu(hi - lo)
    else:
        mine = np.asarray(local_fn(cloud, lo, hi - lo), np.float32).reshape(-1, 3)
    full = all_gather_rows(dist, mine, n, world, device)
    if hasattr(cloud, "normals"):
        cloud.normals = full
    return full


# ---------------------------------------------------------------------------------------------------------
# Sharded VoxelDownSample.  Every rank holds an arbitrary part of the cloud.  One common grid (global
# bounds by an all-reduce), the grid is cut into `world` slabs of consecutive x-indices holding about the same number
# of points (global histogram of the x-index by an all-reduce), ONE all-to-all moves every point to the owner of its
# slab, the owners run the ordinary single-GPU kernel on that common grid, and because the slabs are consecutive in x
# the rank-order concatenation of the results IS the reference's lexicographic output order.  Per-voxel means are
# float64 sums of float32 values (exact, hence order-independent): the result equals the single-GPU result bit for
# bit.  The orchestration below only uses torch tensor ops that exist on CPU and CUDA alike, so the gloo world-2
# test drives exactly this code with a CPU stand-in for the three local library calls (`ops`).
# ---------------------------------------------------------------------------------------------------------
class _GpuVoxelOps:
    """the three local calls of the sharded down-sample on libcupoch_b200.so (CUDA tensors in and out)"""

    def bounds(self, points):
        import numpy as np
        mn, mx = (C.c_float * 3)(), (C.c_float * 3)()
        _lib.check(_lib.lib().cphb_min_max_bound(points.data_ptr(), points.shape[0], mn, mx, None))
        return np.array(mn, np.float32), np.array(mx, np.float32)

    def indices(self, points, voxel, origin):
        import torch
        out = torch.empty((points.shape[0], 3), dtype=torch.int32, device=points.device)
        org = (C.c_float * 3)(*[float(x) for x in origin])
        _lib.check(_lib.lib().cphb_voxel_indices(points.data_ptr(), points.shape[0], float(voxel), org, out.data_ptr(), None))
        _lib.check(_lib.lib().cphb_stream_synchronize(None))
        return out

    def down_sample(self, points, normals, colors, voxel, origin):
        import torch
        m = points.shape[0]
        op = torch.empty((m, 3), dtype=torch.float32, device=points.device)
        on = torch.empty_like(op) if normals is not None else None
        oc = torch.empty_like(op) if colors is not None else None
        org = (C.c_float * 3)(*[float(x) for x in origin])
        k = C.c_size_t(0)
        _lib.check(_lib.lib().cphb_voxel_down_sample_origin(
            points.data_ptr(), normals.data_ptr() if normals is not None else None,
            colors.data_ptr() if colors is not None else None, m, float(voxel), org, op.data_ptr(),
            on.data_ptr() if on is not None else None, oc.data_ptr() if oc is not None else None, C.byref(k), None))
        k = k.value
        return op[:k], (on[:k] if on is not None else None), (oc[:k] if oc is not None else None)


def _exchange_rows(dist, rows, counts, rank, world):
    """all-to-all-v of row blocks: `rows` is sorted by destination, counts[d] rows go to rank d.  NCCL: one
    all_to_all_single; other backends (gloo in the CPU test has no all-to-all): all_gather of the padded buffers."""
    import torch
    cnt = counts.to(torch.int64)
    table = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(table, cnt)
    table = torch.stack(table).cpu()                      # table[s][d] = rows s sends to d
    recv = table[:, rank].tolist()
    if dist.get_backend() == "nccl":
        out = torch.empty((int(sum(recv)),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        dist.all_to_all_single(out, rows.contiguous(), output_split_sizes=[int(x) for x in recv],
                               input_split_sizes=[int(x) for x in cnt.tolist()])
        return out
    biggest = int(table.sum(1).max())
    pad = torch.zeros((max(biggest, 1),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    everyone = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(everyone, pad)
    parts = []
    for s in range(world):
        off = int(table[s, :rank].sum())
        parts.append(everyone[s][off:off + int(table[s, rank])])
    return torch.cat(parts) if parts else rows[:0]


def voxel_down_sample(points, voxel_size, dist, rank, world, normals=None, colors=None, ops=None, gather=True,
                      replicated=False):
    """PointCloud::VoxelDownSample of a cloud spread over `world` ranks.  points / normals / colors: THIS rank's part,
    torch tensors [m, 3] float32 (CUDA with the nccl backend).  Returns (points, normals, colors) tensors of the
    down-sampled cloud: all of it on every rank (gather=True, the reference's lexicographic order) or only this
    rank's slab.  voxel_size <= 0 returns empty tensors like the reference (down_sample.cu:173-176).
    replicated=True: every rank holds the WHOLE cloud (the pyramid of config 5): bounds and histogram are local, each
    rank keeps the points of its own slab, and the only collective is the final all-gather."""
    import numpy as np
    import torch
    ops = ops or _GpuVoxelOps()
    points = points.contiguous()
    normals = None if normals is None else normals.contiguous()
    colors = None if colors is None else colors.contiguous()
    dev = points.device
    empty = points[:0]
    if not voxel_size > 0:
        return empty, (None if normals is None else empty), (None if colors is None else empty)
    v = np.float32(voxel_size)
    m = points.shape[0]
    # 1. the common grid: global bounds, the reference's origin (min_bound - voxel/2, down_sample.cu:180)
    if m:
        mn, mx = ops.bounds(points)
    else:
        mn, mx = np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
    b = torch.from_numpy(np.concatenate([mn, -mx]).astype(np.float32)).to(dev)
    if not replicated:
        dist.all_reduce(b, op=dist.ReduceOp.MIN)
    b = b.cpu().numpy()
    gmin = b[:3].astype(np.float32)
    if not np.isfinite(gmin).all():                       # no point anywhere
        return empty, (None if normals is None else empty), (None if colors is None else empty)
    origin = (gmin - v * np.float32(0.5)).astype(np.float32)
    # 2. slabs of consecutive x-indices with about the same number of points
    kx = ops.indices(points, v, origin)[:, 0].to(torch.int64) if m else torch.zeros(0, dtype=torch.int64, device=dev)
    kmax = torch.tensor([int(kx.max()) if m else 0], dtype=torch.int64, device=dev)
    if not replicated:
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    K = int(kmax.item()) + 1
    hist = torch.bincount(kx, minlength=K)
    if not replicated:
        dist.all_reduce(hist)
    cum = torch.cumsum(hist, 0)
    total = int(cum[-1].item())
    targets = torch.tensor([(total * r) // world for r in range(1, world)], dtype=torch.int64, device=dev)
    # cuts[r-1] = first x-index of rank r's slab: the smallest index whose cumulative count exceeds the target
    cuts = torch.searchsorted(cum, targets, right=True)
    owner = torch.searchsorted(cuts, kx, right=True) if world > 1 else torch.zeros_like(kx)
    cols = [points] + ([normals] if normals is not None else []) + ([colors] if colors is not None else [])
    if replicated:
        # everybody has everything: keep the rows of my own slab, nothing to exchange
        got = torch.cat(cols, 1)[owner == rank]
    else:
        # 3. one exchange: rows sorted by owner (stable: original order inside a destination)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=world)
        rows = torch.cat(cols, 1)[order]
        got = _exchange_rows(dist, rows, counts, rank, world)
    p = got[:, 0:3].contiguous()
    c0 = 3
    nrm = col = None
    if normals is not None:
        nrm = got[:, c0:c0 + 3].contiguous()
        c0 += 3
    if colors is not None:
        col = got[:, c0:c0 + 3].contiguous()
    # 4. the ordinary kernel on the common grid
    if p.shape[0]:
        op, on, oc = ops.down_sample(p, nrm, col, v, origin)
    else:
        op, on, oc = empty, (None if normals is None else empty), (None if colors is None else empty)
    if not gather:
        return op, on, oc
    # 5. slabs are consecutive in x: rank order is the lexicographic order
    outs = []
    for t in (op, on, oc):
        if t is None:
            outs.append(None)
            continue
        n_loc = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
        sizes = [torch.empty_like(n_loc) for _ in range(world)]
        dist.all_gather(sizes, n_loc)
        sizes = [int(s.item()) for s in sizes]
        pad = torch.zeros((max(max(sizes), 1), 3), dtype=t.dtype, device=dev)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        outs.append(torch.cat([parts[r][:sizes[r]] for r in range(world)]))
    return tuple(outs)
