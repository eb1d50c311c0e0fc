#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 (multi-GPU; run under torchrun, one rank per GPU; also works with 1 rank).

  --config 4 : Generalized ICP, 5M -> 5M points (analytic surface + normals -> covariances), source sharded,
               30 iterations, one 32-double exchange per iteration.
  --config 5 : Colored ICP 3-scale pyramid (voxel 0.05/0.025/0.0125, iterations 50/30/14) on a 20M-point
               textured fragment pair over a 4 m x 4 m patch.  Per scale: VoxelDownSample -> EstimateNormals
               (radius 2v, 30) -> colour-gradient init -> ICP.  Pre-processing runs replicated on every
               rank (it needs no collective); the ICP loop shards the down-sampled source.
Rank 0 prints one JSON object.  Sizes can be reduced with --points (stated in the output).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--sharded-prep", action="store_true",
                    help="config 5: shard VoxelDownSample / EstimateNormals over the ranks (cupoch_b200.distributed)")
    args = ap.parse_args()
    import cupoch_b200 as cph
    from cupoch_b200 import _lib
    from cupoch_b200.distributed import destroy_comm, make_comm, shard_range
    from cupoch_b200.testing import datagen
    L = _lib.lib()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", rank))
    _lib.check(L.cphb_set_device(lr))
    dist, comm = None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        comm = make_comm(dist, rank, world, device="cuda", kind=args.comm)
    R, G = cph.registration, cph.geometry
    ev = [L.cphb_event_create() for _ in range(2)]

    def sync():
        L.cphb_stream_synchronize(None)
        if dist is not None:
            dist.barrier()

    def timed(fn):
        sync()
        L.cphb_event_record(ev[0], None)
        out = fn()
        L.cphb_event_record(ev[1], None)
        ms = C.c_float(0)
        L.cphb_event_elapsed_ms(ev[0], ev[1], C.byref(ms))
        if dist is not None:
            import torch
            t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), out
        return ms.value, out

    out = {"config": args.config, "n_gpus": world, "comm": args.comm if world > 1 else None}
    if args.config == 4:
        n = args.points or 5_000_000
        tgt, tn = datagen.surface(n, 11)
        src, sn = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4, attrs=[(tn, True)])
        lo, hi = shard_range(n, rank, world)
        t_pc = G.PointCloud(tgt); t_pc.normals = tn
        s_pc = G.PointCloud(src); s_pc.normals = sn      # full source; the library keeps this rank's spatial block
        shard = (rank, world) if world > 1 else None
        est, crit = R.TransformationEstimationForGeneralizedICP(1e-3), R.ICPConvergenceCriteria(0, 0, 30)
        init_ms, (s_c, t_c) = timed(lambda: (R._with_covariances(s_pc, 1e-3), R._with_covariances(t_pc, 1e-3)))
        run = lambda: R.registration_icp(s_c, t_c, 0.02, np.eye(4), est, crit, comm=comm, return_correspondences=False, shard=shard)
        run()
        ts, loops = [], []
        for _ in range(args.reps):
            ms, res = timed(run)
            ts.append(ms); loops.append(res.loop_ms)
        gt = datagen.gt_transform()
        out.update({"workload": "GICP %d -> %d, 30 iters, r=0.02, eps=1e-3, source sharded x%d" % (n, n, world),
                    "covariance_init_ms": init_ms, "ms_per_registration": float(np.median(ts)),
                    "iters_per_sec": 30e3 / float(np.median(ts)), "loop_iters_per_sec": 30e3 / float(np.median(loops)),
                    "roofline": {"algorithmic_bytes_per_iter": 96 * n, "achieved_gbs": 96 * n / (float(np.median(loops)) / 31) * 1e-6 / world,
                                 "note": "per GPU; 96 B per source point (SURVEY 8d)"},
                    "fitness": res.fitness, "rmse": res.inlier_rmse,
                    "pose_error_vs_ground_truth": float(np.linalg.norm(res.transformation - gt))})
    else:
        n = args.points or 20_000_000
        ext = 4.0
        tgt, tn = datagen.surface(n, 31, extent=ext)
        tc = datagen.texture(tgt, 32, 0.01)
        gt = datagen.gt_transform((0.0, 0.0, 2.0), (0.01, 0.0, 0.0))
        src, sc = datagen.make_source(tgt, gt, 33, 34, 2e-4, attrs=[(tc, False)])
        T = np.eye(4, dtype=np.float32)
        t_full = G.PointCloud(tgt); t_full.colors = tc
        s_full = G.PointCloud(src); s_full.colors = sc
        stages = []

        def pyramid():
            T = np.eye(4, dtype=np.float32)
            st = []
            for v, iters in ((0.05, 50), (0.025, 30), (0.0125, 14)):
                t0 = time.perf_counter()
                if args.sharded_prep and world > 1:
                    # every rank down-samples its own slab of the (replicated) clouds, one all-gather;
                    # normals by blocks of queries against the replicated index, one all-gather
                    import torch
                    from cupoch_b200 import distributed as D

                    def down(pc):
                        tp = torch.as_tensor(pc.points, device="cuda")
                        tcol = torch.as_tensor(pc.colors, device="cuda")
                        p_, _, c_ = D.voxel_down_sample(tp, v, dist, rank, world, colors=tcol, replicated=True)
                        o = G.PointCloud(p_)
                        o.colors = c_
                        return o
                    td, sd = down(t_full), down(s_full)
                    L.cphb_stream_synchronize(None); t1 = time.perf_counter()
                    D.estimate_normals(td, G.KDTreeSearchParamRadius(2 * v, 30), dist, rank, world, device="cuda")
                    D.estimate_normals(sd, G.KDTreeSearchParamRadius(2 * v, 30), dist, rank, world, device="cuda")
                else:
                    td, sd = t_full.voxel_down_sample(v), s_full.voxel_down_sample(v)
                    L.cphb_stream_synchronize(None); t1 = time.perf_counter()
                    td.estimate_normals(G.KDTreeSearchParamRadius(2 * v, 30))
                    sd.estimate_normals(G.KDTreeSearchParamRadius(2 * v, 30))
                L.cphb_stream_synchronize(None); t2 = time.perf_counter()
                ns = len(sd)
                res = R.registration_colored_icp(sd, td, v, T, R.ICPConvergenceCriteria(1e-6, 1e-6, iters), comm=comm,
                                                 return_correspondences=False, shard=(rank, world) if world > 1 else None)
                L.cphb_stream_synchronize(None); t3 = time.perf_counter()
                T = res.transformation
                st.append({"voxel": v, "n_src": ns, "n_tgt": len(td), "voxel_ms": 1e3 * (t1 - t0), "normals_ms": 1e3 * (t2 - t1),
                           "icp_ms": 1e3 * (t3 - t2), "iterations": res.iterations, "loop_ms": res.loop_ms,
                           "fitness": res.fitness, "rmse": res.inlier_rmse})
            return T, st
        pyramid()
        ms, (T, stages) = timed(pyramid)
        out.update({"workload": "Colored ICP 3-scale pyramid on %d-point pair (4m x 4m patch), source sharded x%d for the ICP loop"
                                % (n, world),
                    "end_to_end_ms": ms, "stages": stages,
                    "pose_error_vs_ground_truth": float(np.linalg.norm(T - gt))})
    if rank == 0:
        print(json.dumps(out))
    if comm is not None:
        sync()
        destroy_comm(comm)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
