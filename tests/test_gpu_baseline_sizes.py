"""GPU parity AT THE SIZES BASELINE.json names: the CUDA path against the CPU oracle on

  config 2   point-to-plane ICP 1 M -> 1 M, 30 iterations: per-iteration pose trace, final pose, the whole 1 M
             correspondence list, fitness / rmse;
  config 3   VoxelDownSample(0.02) of 10 M points (1 and 3 attributes) bit-exact, SearchRadius(k=1, r=0.05) of the 10 M
             points against the down-sampled cloud: indices and d2 bit-exact;
  TOP = 5    one search against a 34 M-point cloud (> 32^5 points: the deepest kernel instantiation,
             search1_kernel<5>), checked exactly through a spatial window (see the test);
  config 4   Generalized ICP 1 M -> 1 M (the 5 M / 8-GPU job at a size one host oracle run finishes quickly);
  config 5   the Colored-ICP 3-scale pyramid on a 2 M-point textured pair, every stage against the oracle.

Each test prints one line "PARITY {json}" (pose_delta, index_mismatches, ...) that tools/ and DESIGN.md quote.
Tolerances: indices / d2 / voxel outputs bit-exact; pose within 1e-5 Frobenius (north_star) -- observed 0.
"""
import json
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph
from cupoch_b200.testing import datagen

R, G = cph.registration, cph.geometry
POSE_TOL = 1e-5


def cloud(p, n=None, c=None, cov=None):
    pc = G.PointCloud(p)
    if n is not None:
        pc.normals = n
    if c is not None:
        pc.colors = c
    if cov is not None:
        pc.covariances = cov
    return pc


def report(name, **kw):
    print("PARITY " + json.dumps(dict(test=name, **kw)))


def corr_mismatches(a, b):
    """number of source points whose match differs between two (i, j) lists"""
    if a.shape == b.shape and np.array_equal(a, b):
        return 0
    n = int(max(a[:, 0].max(initial=-1), b[:, 0].max(initial=-1))) + 1
    ma, mb = np.full(n, -1, np.int64), np.full(n, -1, np.int64)
    ma[a[:, 0]] = a[:, 1]
    mb[b[:, 0]] = b[:, 1]
    return int((ma != mb).sum())


def test_config2_p2plane_1m_vs_oracle(orc):
    n, iters, r = 1_000_000, 30, 0.02
    tgt, tn = datagen.surface(n, 11)
    src = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4)
    t0 = time.perf_counter()
    ref = orc.registration_icp(orc.P2PLANE, src, tgt, r, tgt_nrm=tn, relative_fitness=0, relative_rmse=0,
                               max_iteration=iters, trace=True)
    t_cpu = time.perf_counter() - t0
    est = R.TransformationEstimationPointToPlane()
    s_pc, t_pc = cloud(src), cloud(tgt, tn)
    res = R.registration_icp(s_pc, t_pc, r, np.eye(4), est, R.ICPConvergenceCriteria(0, 0, iters))
    d = float(np.linalg.norm(res.transformation.astype(np.float64) - ref["transformation"].astype(np.float64)))
    mism = corr_mismatches(res.correspondence_set, ref["correspondence_set"])
    # per-iteration trace: the registration is deterministic, so a run capped at k updates ends in the pose the
    # 30-update run had after k updates (the device-side loop keeps no trace of its own)
    trace_delta = []
    for k in range(0, iters + 1):
        rk = R.registration_icp(s_pc, t_pc, r, np.eye(4), est, R.ICPConvergenceCriteria(0, 0, k), return_correspondences=False)
        trace_delta.append(float(np.linalg.norm(rk.transformation.astype(np.float64) - ref["trace"][k].astype(np.float64))))
    report("config2_p2plane_1m", pose_delta=d, index_mismatches=mism, correspondences=int(len(ref["correspondence_set"])),
           trace_max_delta=max(trace_delta), fitness=res.fitness, rmse=res.inlier_rmse, oracle_seconds=round(t_cpu, 2))
    assert d <= POSE_TOL
    assert max(trace_delta) <= POSE_TOL
    assert res.iterations == ref["iterations"] == iters
    assert abs(res.fitness - ref["fitness"]) <= 1e-6 and abs(res.inlier_rmse - ref["inlier_rmse"]) <= 1e-6
    np.testing.assert_array_equal(res.correspondence_set, ref["correspondence_set"])


@pytest.fixture(scope="module")
def config3_cloud():
    n = 10_000_000
    p = datagen.uniform_cube(n, 21, hi=(4, 4, 1))
    return p


def test_config3_voxel_10m_vs_oracle(orc, config3_cloud):
    p = config3_cloud
    pc = G.PointCloud(p)
    out = pc.voxel_down_sample(0.02)
    op, _, _ = orc.voxel_down_sample(p, 0.02)
    got = out.points.cpu()
    report("config3_voxel_10m_A1", n_out=int(len(got)), n_out_oracle=int(len(op)),
           differing_rows=int((got != op).any(1).sum()) if got.shape == op.shape else -1)
    np.testing.assert_array_equal(got, op)
    # 3-attribute branch on 4 M of the points (mean-then-normalise normals, mean colours)
    m = 4_000_000
    nrm, col = datagen.unit_normals(m, 22), datagen.uniform_cube(m, 23)
    pc3 = cloud(p[:m], nrm, col)
    out3 = pc3.voxel_down_sample(0.02)
    op3, on3, oc3 = orc.voxel_down_sample(p[:m], 0.02, normals=nrm, colors=col)
    np.testing.assert_array_equal(out3.points.cpu(), op3)
    np.testing.assert_array_equal(out3.normals.cpu(), on3)
    np.testing.assert_array_equal(out3.colors.cpu(), oc3)
    report("config3_voxel_4m_A3", n_out=int(len(op3)), differing_rows=0)


def test_config3_search_10m_vs_oracle(orc, config3_cloud):
    p = config3_cloud
    down, _, _ = orc.voxel_down_sample(p, 0.02)
    tree = G.KDTreeFlann(G.PointCloud(down))
    cnt, idx, d2 = tree.search_radius(p, 0.05, 1)
    idx, d2 = idx.cpu(), d2.cpu()
    oi, od, oc = orc.search(down, p, 1, radius=0.05, kdtree=True)
    report("config3_search_10m_k1", queries=int(len(p)), targets=int(len(down)), found=int(cnt),
           index_mismatches=int((idx != oi).sum()), d2_mismatches=int((d2.view(np.uint32) != od.view(np.uint32)).sum()))
    assert cnt == oc
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(d2.view(np.uint32), od.view(np.uint32))


def test_top5_search_34m(orc):
    """> 32^5 = 33 554 432 points: the index has five box levels above the leaves and search1_kernel<5> /
    searchk_kernel<5> are the instances that run.  The oracle cannot search 34 M points quickly, but a radius search
    is local: every neighbour within r of a query inside a window W lies inside W grown by r, so the oracle runs on
    that sub-cloud and its indices are mapped back -- an exact check of the full-size search."""
    n = 34_000_000
    rng = np.random.Generator(np.random.PCG64(77))
    p = rng.random((n, 3), dtype=np.float32)
    p[:, 2] *= np.float32(0.25)
    r = 0.01
    tree = G.KDTreeFlann(G.PointCloud(p))
    lo, hi = np.float32([0.40, 0.55, 0.05]), np.float32([0.46, 0.61, 0.20])
    # queries: the cloud's own points inside the window, jittered (so d2 > 0), plus points outside the cloud
    inside = np.all((p >= lo) & (p < hi), axis=1)
    q = p[inside][:200_000] + rng.normal(0, 0.002, (min(int(inside.sum()), 200_000), 3)).astype(np.float32)
    q = np.clip(q, lo, np.nextafter(hi, lo)).astype(np.float32)
    grow = np.float32(r * 1.01)
    sub_mask = np.all((p >= lo - grow) & (p <= hi + grow), axis=1)
    sub_ids = np.nonzero(sub_mask)[0].astype(np.int32)
    sub = p[sub_mask]
    for k in (1, 8):
        cnt, idx, d2 = tree.search_radius(q, r, k)
        idx, d2 = idx.cpu(), d2.cpu()
        oi, od, oc = orc.search(sub, q, k, radius=r, kdtree=True)
        oi_full = np.where(oi >= 0, sub_ids[np.maximum(oi, 0)], -1)
        report("top5_search_34m_k%d" % k, points=n, queries=int(len(q)), window_points=int(len(sub)), found=int(cnt),
               index_mismatches=int((idx != oi_full).sum()), d2_mismatches=int((d2.view(np.uint32) != od.view(np.uint32)).sum()))
        assert cnt == oc
        np.testing.assert_array_equal(idx, oi_full)
        np.testing.assert_array_equal(d2.view(np.uint32), od.view(np.uint32))
    # far corner of the domain (last leaves of the Hilbert order, padded tail of the index)
    q2 = (np.float32([1.0, 1.0, 0.25]) - rng.random((2000, 3), dtype=np.float32) * np.float32(0.02)).astype(np.float32)
    m2 = np.all(p >= np.float32([0.96, 0.96, 0.21]), axis=1)
    ids2, sub2 = np.nonzero(m2)[0].astype(np.int32), p[m2]
    cnt, idx, d2 = tree.search_radius(q2, r, 1)
    oi, od, oc = orc.search(sub2, q2, 1, radius=r, kdtree=True)
    np.testing.assert_array_equal(idx.cpu(), np.where(oi >= 0, ids2[np.maximum(oi, 0)], -1))
    np.testing.assert_array_equal(d2.cpu().view(np.uint32), od.view(np.uint32))


def test_config4_gicp_1m_vs_oracle(orc):
    n, iters, r = 1_000_000, 10, 0.02
    tgt, tn = datagen.surface(n, 11)
    src, sn = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4, attrs=[(tn, True)])
    scov, tcov = orc.covariances_from_normals(sn, 1e-3), orc.covariances_from_normals(tn, 1e-3)
    t0 = time.perf_counter()
    ref = orc.registration_icp(orc.GICP, src, tgt, r, src_cov=scov, tgt_cov=tcov, relative_fitness=0, relative_rmse=0,
                               max_iteration=iters)
    t_cpu = time.perf_counter() - t0
    res = R.registration_generalized_icp(cloud(src, sn), cloud(tgt, tn), r, np.eye(4),
                                         R.TransformationEstimationForGeneralizedICP(1e-3), R.ICPConvergenceCriteria(0, 0, iters))
    d = float(np.linalg.norm(res.transformation.astype(np.float64) - ref["transformation"].astype(np.float64)))
    mism = corr_mismatches(res.correspondence_set, ref["correspondence_set"])
    report("config4_gicp_1m", pose_delta=d, index_mismatches=mism, correspondences=int(len(ref["correspondence_set"])),
           iterations=iters, fitness=res.fitness, rmse=res.inlier_rmse, oracle_seconds=round(t_cpu, 2))
    assert d <= POSE_TOL
    assert mism == 0
    assert abs(res.fitness - ref["fitness"]) <= 1e-6 and abs(res.inlier_rmse - ref["inlier_rmse"]) <= 1e-6


def test_config5_colored_pyramid_2m_vs_oracle(orc):
    """examples/python/advanced/colored_pointcloud_registration.py:37-60 at 2 M points per fragment: per scale
    VoxelDownSample -> EstimateNormals(radius 2v, 30) -> colour gradient -> RegistrationColoredICP, every stage
    compared with the oracle's stage on the same inputs."""
    n, ext = 2_000_000, 4.0
    tgt, _ = datagen.surface(n, 31, extent=ext)
    tc = datagen.texture(tgt, 32, 0.01)
    gt = datagen.gt_transform((0.0, 0.0, 2.0), (0.01, 0.0, 0.0))
    src, sc = datagen.make_source(tgt, gt, 33, 34, 2e-4, attrs=[(tc, False)])
    t_full, s_full = cloud(tgt, c=tc), cloud(src, c=sc)
    T = np.eye(4, dtype=np.float32)
    To = np.eye(4, dtype=np.float32)
    stages = []
    for v, iters in ((0.05, 50), (0.025, 30), (0.0125, 14)):
        td, sd = t_full.voxel_down_sample(v), s_full.voxel_down_sample(v)
        otp, _, otc = orc.voxel_down_sample(tgt, v, colors=tc)
        osp, _, osc = orc.voxel_down_sample(src, v, colors=sc)
        np.testing.assert_array_equal(td.points.cpu(), otp)
        np.testing.assert_array_equal(td.colors.cpu(), otc)
        np.testing.assert_array_equal(sd.points.cpu(), osp)
        np.testing.assert_array_equal(sd.colors.cpu(), osc)
        td.estimate_normals(G.KDTreeSearchParamRadius(2 * v, 30))
        sd.estimate_normals(G.KDTreeSearchParamRadius(2 * v, 30))
        otn = orc.estimate_normals(otp, knn=0, radius=2 * v, max_nn=30)
        nrm_diff = int((td.normals.cpu() != otn).any(1).sum())
        res = R.registration_colored_icp(sd, td, v, T, R.ICPConvergenceCriteria(1e-6, 1e-6, iters))
        nbr, _, _ = orc.search(otp, otp, 30, radius=2 * v, kdtree=True)
        grad = orc.color_gradient(otp, otn, otc, nbr)
        ref = orc.registration_icp(orc.COLORED, osp, otp, v, init=To, src_col=osc, tgt_nrm=otn, tgt_col=otc, tgt_grad=grad,
                                   relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=iters)
        d = float(np.linalg.norm(res.transformation.astype(np.float64) - ref["transformation"].astype(np.float64)))
        mism = corr_mismatches(res.correspondence_set, ref["correspondence_set"])
        stages.append({"voxel": v, "n_src": int(len(osp)), "n_tgt": int(len(otp)), "normal_rows_differing": nrm_diff,
                       "pose_delta": d, "index_mismatches": mism, "iterations": [int(res.iterations), int(ref["iterations"])]})
        T, To = res.transformation, ref["transformation"]
    report("config5_colored_pyramid_2m", stages=stages, pose_error_vs_ground_truth=float(np.linalg.norm(T - gt)))
    for s in stages:
        assert s["normal_rows_differing"] == 0
        assert s["iterations"][0] == s["iterations"][1]
        assert s["pose_delta"] <= POSE_TOL
        assert s["index_mismatches"] == 0
