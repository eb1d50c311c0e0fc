"""CPU: the oracle's restatements of the reducer's other users (ComputeJTJandJTr / ComputeWeightedJTJandJTr on rows,
KabschWeighted), FPFH and DBSCAN against independent float64 numpy / scipy computations."""
import numpy as np
import pytest

from cupoch_b200.testing import datagen


def _rows(n, num_j, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, num_j, 6)).astype(np.float32), (0.1 * rng.standard_normal((n, num_j))).astype(np.float32)


def _unpack(S):
    JTJ = np.zeros((6, 6))
    p = 0
    for a in range(6):
        for b in range(a, 6):
            JTJ[a, b] = JTJ[b, a] = S[p]
            p += 1
    return JTJ, S[21:27], S[27]


def test_jtj_rows_against_numpy(orc):
    J, r = _rows(5000, 2, 1)
    JTJ, JTr, r2 = _unpack(orc.jtj_rows(J, r))
    Jd, rd = J.reshape(-1, 6).astype(np.float64), r.reshape(-1).astype(np.float64)
    np.testing.assert_allclose(JTJ, Jd.T @ Jd, rtol=2e-6)
    np.testing.assert_allclose(JTr, Jd.T @ rd, rtol=2e-5, atol=1e-4)
    assert abs(r2 - rd @ rd) < 1e-5 * r2


def test_weighted_jtj_rows_against_numpy(orc):
    J, r = _rows(4000, 2, 2)
    sigma2, nu = 0.05, 5.0
    S, w_sum = orc.weighted_jtj_rows(J, r, sigma2, nu)
    Jd, rd = J.astype(np.float64), r.astype(np.float64)
    r2 = (rd ** 2).sum(1)
    ws = (r2 * (nu + 1.0) / (nu + r2 / sigma2)).sum()
    assert abs(w_sum - ws) < 1e-5 * ws
    w = (nu + 1.0) / (nu + r2 / ws)
    JTJ = np.einsum("i,ija,ijb->ab", w, Jd, Jd)
    JTr = np.einsum("i,ija,ij->a", w, Jd, rd)
    g, b, rr = _unpack(S)
    np.testing.assert_allclose(g, JTJ, rtol=1e-5)
    np.testing.assert_allclose(b, JTr, rtol=1e-4, atol=1e-3)
    assert abs(rr - (w * r2).sum()) < 1e-5 * rr


def test_kabsch_weighted_recovers_a_rigid_motion(orc):
    rng = np.random.default_rng(3)
    m = rng.random((3000, 3)).astype(np.float32)
    T = datagen.gt_transform((10.0, -20.0, 30.0), (0.3, -0.2, 0.1))
    t = (m.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    w = rng.random(3000).astype(np.float32) + 0.1
    got = orc.kabsch_weighted(m, t, w)
    np.testing.assert_allclose(got, T, atol=2e-5)
    # uniform weights == plain Kabsch on all pairs
    got1 = orc.kabsch_weighted(m, t, np.ones(3000, np.float32))
    np.testing.assert_allclose(got1, T, atol=2e-5)


def test_det_atan2_is_correctly_rounded(orc):
    import ctypes as C
    L = orc.lib()
    L.orc_det_atan2f.restype = C.c_float
    L.orc_det_atan2f.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(0)
    y = rng.standard_normal(20000).astype(np.float32)
    x = rng.standard_normal(20000).astype(np.float32)
    y[:6] = [0, 0, 1, -1, 1e-30, -1e-30]
    x[:6] = [1, -1, 0, 0, 1, -1]
    got = np.array([L.orc_det_atan2f(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    np.testing.assert_array_equal(got, np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32))


def test_fpfh_properties(orc):
    p, n = datagen.surface(3000, 5)
    f = orc.compute_fpfh_feature(p, n, knn=12)
    assert f.shape == (3000, 33) and np.isfinite(f).all() and (f >= 0).all()
    # each 11-bin block: the point's own SPFH (sums to 100) + the re-normalised neighbour sum (100)
    np.testing.assert_allclose(f.reshape(-1, 3, 11).sum(2), 200.0, rtol=1e-4)
    fr = orc.compute_fpfh_feature(p, n, radius=0.05, max_nn=20)
    assert np.isfinite(fr).all()
    # rigid motions leave the descriptor (nearly) unchanged
    T = datagen.gt_transform((20.0, 10.0, -30.0), (0.5, 0.1, -0.2))
    p2 = (p.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    n2 = (n.astype(np.float64) @ T[:3, :3].T).astype(np.float32)
    f2 = orc.compute_fpfh_feature(p2, n2, knn=12)
    assert np.median(np.abs(f2 - f).sum(1)) < 5.0


def test_dbscan_against_connected_components(orc):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(7)
    blobs = [rng.normal(c, 0.1, (400, 3)) for c in ((0, 0, 0), (1, 0, 0), (0, 1, 0.5), (1, 1, 1))]
    noise = rng.uniform(-1, 2, (150, 3))
    pts = np.concatenate(blobs + [noise]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    eps, min_points = 0.06, 6
    labels, k = orc.cluster_dbscan(pts, eps, min_points, max_edges=100)
    # textbook DBSCAN on the same (untruncated, here: lists are never full) graph: components of the core points
    tree = cKDTree(pts.astype(np.float64))
    nb = tree.query_ball_point(pts.astype(np.float64), eps * (1 - 1e-6))
    deg = np.array([len(x) - 1 for x in nb])
    core = deg >= min_points
    assert max(deg) < 100
    rows, cols = [], []
    for i in np.nonzero(core)[0]:
        for j in nb[i]:
            if core[j]:
                rows.append(i)
                cols.append(j)
    ncomp, comp = connected_components(coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(len(pts),) * 2), directed=False)
    # every core point's label identifies its component; cluster ids are consecutive in seed order
    lab_core = labels[core]
    assert (lab_core >= 0).all()
    m = {}
    for l, c in zip(lab_core, comp[core]):
        assert m.setdefault(l, c) == c
    assert len(set(m.values())) == len(m) == k
    assert sorted(m.keys()) == list(range(k))
    # noise: not core and no core neighbour
    for i in np.nonzero(labels < 0)[0]:
        assert not core[i] and not any(core[j] for j in nb[i])
    assert (labels[~core] >= 0).sum() > 0          # there are border points
