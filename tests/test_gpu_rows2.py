"""GPU parity of the rows added in round 2: ComputeJTJandJTr / ComputeWeightedJTJandJTr on explicit
rows, KabschWeighted, ComputeFPFHFeature, ClusterDBSCAN -- C ABI through the Python mirror vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph
from cupoch_b200.testing import datagen


def _rows(n, num_j, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, num_j, 6)).astype(np.float32), (0.1 * rng.standard_normal((n, num_j))).astype(np.float32)


def _sums_equal(S_gpu, S_orc):
    # float64 sums of the same float32 values in different orders: identical after the float32 rounding (<= 1 ulp on a
    # rounding boundary)
    a, b = np.asarray(S_gpu, np.float64), np.asarray(S_orc, np.float64)
    np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-12 * np.abs(b).max())
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    assert (np.abs(a32.view(np.int32) - b32.view(np.int32)) <= 1).all()


def _pack(JTJ, JTr, r2):
    out = [JTJ[a, b] for a in range(6) for b in range(a, 6)] + list(JTr) + [r2]
    return np.array(out, np.float64)


@pytest.mark.parametrize("n,num_j", [(1, 1), (777, 1), (100_000, 2), (300_000, 3)])
def test_compute_jtj_jtr(orc, n, num_j):
    J, r = _rows(n, num_j, n)
    JTJ, JTr, r2 = cph.utility.compute_jtj_jtr(J, r)
    S = orc.jtj_rows(J, r)
    ref = np.concatenate([S[:27], S[27:28]]).astype(np.float32).astype(np.float64)
    got = _pack(JTJ, JTr, r2)
    assert (np.abs(got.astype(np.float32).view(np.int32) - ref.astype(np.float32).view(np.int32)) <= 1).all()


def test_compute_weighted_jtj_jtr(orc):
    J, r = _rows(200_000, 2, 5)
    JTJ, JTr, r2, w_sum = cph.utility.compute_weighted_jtj_jtr(J, r, 0.05, 5.0)
    S, ws = orc.weighted_jtj_rows(J, r, 0.05, 5.0)
    assert np.float32(w_sum) == np.float32(ws)
    ref = np.concatenate([S[:27], S[27:28]]).astype(np.float32)
    got = _pack(JTJ, JTr, r2).astype(np.float32)
    assert (np.abs(got.view(np.int32) - ref.view(np.int32)) <= 1).all()


def test_kabsch_weighted(orc):
    rng = np.random.default_rng(3)
    m = rng.random((50_000, 3)).astype(np.float32)
    T = datagen.gt_transform((10.0, -20.0, 30.0), (0.3, -0.2, 0.1))
    t = (m.astype(np.float64) @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 1e-3, m.shape)).astype(np.float32)
    w = rng.random(len(m)).astype(np.float32) + 0.1
    got = cph.registration.kabsch_weighted(m, t, w)
    ref = orc.kabsch_weighted(m, t, w)
    np.testing.assert_allclose(got, ref, atol=1e-6, rtol=0)
    np.testing.assert_allclose(got, T, atol=1e-3)


@pytest.mark.parametrize("param", [("knn", 15), ("radius", 0.04, 25)])
def test_fpfh_bit_exact(orc, param):
    p, n = datagen.surface(40_000, 5)
    pc = cph.geometry.PointCloud(p)
    pc.normals = n
    if param[0] == "knn":
        f = cph.registration.compute_fpfh_feature(pc, cph.geometry.KDTreeSearchParamKNN(param[1])).cpu()
        ref = orc.compute_fpfh_feature(p, n, knn=param[1])
    else:
        f = cph.registration.compute_fpfh_feature(pc, cph.geometry.KDTreeSearchParamRadius(param[1], param[2])).cpu()
        ref = orc.compute_fpfh_feature(p, n, radius=param[1], max_nn=param[2])
    assert f.shape == (len(p), 33)
    np.testing.assert_array_equal(f, ref)


def test_fpfh_needs_normals():
    pc = cph.geometry.PointCloud(datagen.uniform_cube(100, 1))
    with pytest.raises(RuntimeError):
        cph.registration.compute_fpfh_feature(pc, cph.geometry.KDTreeSearchParamKNN(10))


def _blobs(seed, n_blob=4000, n_noise=1500):
    rng = np.random.default_rng(seed)
    blobs = [rng.normal(c, 0.1, (n_blob, 3)) for c in ((0, 0, 0), (1, 0, 0), (0, 1, 0.5), (1, 1, 1), (2, 2, 0))]
    pts = np.concatenate(blobs + [rng.uniform(-1, 3, (n_noise, 3))]).astype(np.float32)
    return pts[rng.permutation(len(pts))]


@pytest.mark.parametrize("eps,min_points,max_edges", [(0.03, 6, 100), (0.05, 10, 20), (0.02, 3, 5)])
def test_dbscan_labels_equal_oracle(orc, eps, min_points, max_edges):
    pts = _blobs(7)
    pc = cph.geometry.PointCloud(pts)
    labels = pc.cluster_dbscan(eps, min_points, False, max_edges).cpu()
    ref, k = orc.cluster_dbscan(pts, eps, min_points, max_edges)
    assert pc.last_cluster_count == k
    np.testing.assert_array_equal(labels, ref)
    assert k >= 5 and (labels < 0).any()


def test_dbscan_surface_200k(orc):
    # one connected sheet: a single cluster; BFS depth in the hundreds
    p, _ = datagen.surface(200_000, 9)
    pc = cph.geometry.PointCloud(p)
    labels = pc.cluster_dbscan(0.008, 4, False, 30).cpu()
    ref, k = orc.cluster_dbscan(p, 0.008, 4, 30)
    np.testing.assert_array_equal(labels, ref)
