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


def estimate_normals(cloud, search_param, dist, rank, world, device=None, local_fn=None):
    """PointCloud::EstimateNormals over `world` ranks (replicate the index, shard the queries, no
    collective in the search; one all-gather of the per-point outputs).  Every rank holds the full cloud, estimates
    the normals of its contiguous block with cphb_estimate_normals_range and all-gathers the blocks: the result is
    bit-identical to the single-GPU estimate.  `local_fn(points, first, count) -> [count, 3]` replaces the library
    call in the gloo/CPU test of this orchestration."""
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
        mine = out.cpu(hi - lo)
    else:
        mine = np.asarray(local_fn(cloud, lo, hi - lo), np.float32).reshape(-1, 3)
    full = all_gather_rows(dist, mine, n, world, device)
    if hasattr(cloud, "normals"):
        cloud.normals = full
    return full
