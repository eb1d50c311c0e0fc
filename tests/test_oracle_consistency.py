"""CPU: internal consistency of the oracle (the checker must be trustworthy before it checks the GPU path):
kd-tree == brute force (incl. ties and the strict radius), ICP recovers a known pose for every estimator,
voxel means against numpy, Kabsch quirk, LDLT solve against numpy, transform/covariance algebra."""
import numpy as np
import pytest

from cupoch_b200.testing import datagen


def test_kdtree_equals_bruteforce(orc):
    rng = np.random.default_rng(0)
    tgt = rng.random((3000, 3), dtype=np.float32)
    tgt = np.concatenate([tgt, tgt[:200], np.round(tgt[:300], 1)])       # duplicates and lattice ties
    qry = np.concatenate([rng.random((500, 3), dtype=np.float32), np.round(rng.random((200, 3)), 1).astype(np.float32)])
    for k, r in ((1, 0.05), (1, -1.0), (7, 0.1), (30, -1.0), (100, 0.2)):
        a = orc.search(tgt, qry, k, radius=r, kdtree=False)
        b = orc.search(tgt, qry, k, radius=r, kdtree=True)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert a[2] == b[2]
        d2 = a[1]
        fin = np.isfinite(d2)
        np.testing.assert_array_equal(d2, np.sort(d2, axis=1))            # ascending, unfilled (+inf) last
        if r > 0:
            assert (d2[fin] < np.float32(r) * np.float32(r)).all()        # strict


@pytest.mark.parametrize("kind", ["p2p", "p2plane", "symmetric", "gicp", "colored"])
def test_icp_recovers_known_pose(orc, kind):
    n = 4000
    tgt, tn = datagen.surface(n, 11)
    tc = datagen.texture(tgt)
    gt = datagen.gt_transform((-0.6, 0.9, 1.2), (0.006, -0.003, 0.004))
    src, sn, sc = datagen.make_source(tgt, gt, 13, 14, 0.0, attrs=[(tn, True), (tc, False)])
    kw = dict(relative_fitness=1e-9, relative_rmse=1e-9, max_iteration=40)
    r = 0.06
    if kind == "p2p":
        res = orc.registration_icp(orc.P2P, src, tgt, r, **kw)
    elif kind == "p2plane":
        res = orc.registration_icp(orc.P2PLANE, src, tgt, r, tgt_nrm=tn, **kw)
    elif kind == "symmetric":
        res = orc.registration_icp(orc.SYMMETRIC, src, tgt, r, src_nrm=sn, tgt_nrm=tn, **kw)
    elif kind == "gicp":
        res = orc.registration_icp(orc.GICP, src, tgt, r, src_cov=orc.covariances_from_normals(sn),
                                   tgt_cov=orc.covariances_from_normals(tn), **kw)
    else:
        nbr, _, _ = orc.search(tgt, tgt, 30, radius=2 * r, kdtree=True)
        res = orc.registration_icp(orc.COLORED, src, tgt, r, src_col=sc, tgt_nrm=tn, tgt_col=tc,
                                   tgt_grad=orc.color_gradient(tgt, tn, tc, nbr), **kw)
    assert res["fitness"] > 0.99
    assert np.linalg.norm(res["transformation"] - gt) < (5e-3 if kind in ("p2p", "colored") else 5e-4)


def test_solve_jtj_matches_numpy(orc):
    rng = np.random.default_rng(1)
    J = rng.standard_normal((500, 6))
    r = rng.standard_normal(500) * 0.01
    A, b = J.T @ J, J.T @ r
    u = np.array([A[i, j] for i in range(6) for j in range(i, 6)], np.float32)
    ok, T = orc.solve_jtj(u, b.astype(np.float32))
    x = np.linalg.solve(A, -b)
    th = np.linalg.norm(x[:3])
    k = x[:3] / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    assert ok
    np.testing.assert_allclose(T[:3, :3], R, atol=2e-6)
    np.testing.assert_allclose(T[:3, 3], x[3:], atol=2e-6)
    ok, T = orc.solve_jtj(np.zeros(21, np.float32), np.zeros(6, np.float32))    # singular: det check fails
    assert not ok and np.array_equal(T, np.eye(4, dtype=np.float32))


def test_kabsch_divides_by_model_size(orc):
    """kabsch.cu:76-78,107: sums are divided by model.size(), not by the number of correspondences; with all
    points matched this is the ordinary Kabsch, with a subset it is the reference's biased variant."""
    rng = np.random.default_rng(2)
    p = rng.random((200, 3)).astype(np.float32)
    T = datagen.gt_transform((3, -2, 4), (0.1, 0.2, -0.1)).astype(np.float32)
    q = orc.transform_points(p, T)
    full = np.stack([np.arange(200)] * 2, 1).astype(np.int32)
    Tk, _ = orc.kabsch(p, q, full)
    np.testing.assert_allclose(Tk, T, atol=2e-5)
    Th, _ = orc.kabsch(p, q, full[:100])
    assert np.linalg.norm(Th - T) > 1e-3      # the quirk is real ...
    Th2, _ = orc.kabsch(p, q, full[:100], n_model=100)
    np.testing.assert_allclose(Th2, T, atol=5e-5)   # ... and disappears when dividing by the match count


def test_voxel_means_against_numpy(orc):
    p = datagen.uniform_cube(20000, 3)
    c = datagen.uniform_cube(20000, 4)
    op, _, oc = orc.voxel_down_sample(p, 0.1, None, c)
    org = p.min(0) - np.float32(0.05)
    key = np.floor((p - org) / np.float32(0.1)).astype(np.int64)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, (np.diff(ks, axis=0) != 0).any(1)])
    means = np.add.reduceat(p[order].astype(np.float64), starts) / np.diff(np.r_[starts, len(p)])[:, None]
    assert len(op) == len(starts)
    np.testing.assert_allclose(op, means, rtol=2e-6)
    assert (op.min(0) >= p.min(0) - 1e-6).all() and (op.max(0) <= p.max(0) + 1e-6).all()


def test_covariance_rotation_algebra(orc):
    nrm = datagen.unit_normals(100, 1)
    cov = orc.covariances_from_normals(nrm, 1e-3)
    np.testing.assert_allclose(np.einsum("nij,nj->ni", cov, nrm), 1e-3 * nrm, atol=2e-6)   # eps along the normal
    T = datagen.gt_transform((10, 20, 30), (1, 2, 3)).astype(np.float32)
    rot = orc.rotate_covariances(cov, T)
    R = T[:3, :3].astype(np.float64)
    np.testing.assert_allclose(rot, np.einsum("ij,njk,lk->nil", R, cov, R), atol=2e-6)


def test_deterministic_acos_cos_are_correctly_rounded(orc):
    """FastEigen3x3's acos / cos are a float64-polynomial specification shared by oracle and kernel (oracle.c
    "deterministic acos / cos").  It must also be an accurate acosf / cosf: here, equal to the correctly rounded value."""
    import ctypes as C
    L = orc.lib()
    L.orc_det_acosf.restype = C.c_float
    L.orc_det_acosf.argtypes = [C.c_float]
    L.orc_det_cosf.restype = C.c_float
    L.orc_det_cosf.argtypes = [C.c_float]
    rng = np.random.default_rng(0)
    f32 = np.float32
    xs = np.concatenate([rng.uniform(-1, 1, 20000).astype(f32),
                         f32([-1, 1, 0, 0.5, -0.5, np.nextafter(f32(0.5), f32(1)), np.nextafter(f32(-0.5), f32(-1)),
                              0.99999994, -0.99999994, 1e-20, -1e-20])])
    got = np.array([L.orc_det_acosf(float(x)) for x in xs], f32)
    np.testing.assert_array_equal(got, np.arccos(xs.astype(np.float64)).astype(f32))
    ys = np.concatenate([rng.uniform(0, 2 * np.pi, 20000).astype(f32),
                         f32([0, np.pi / 4, np.pi / 2, 3 * np.pi / 4, np.pi, 2 * np.pi, 3.1415929, 2.0943952, 1.0471976])])
    got = np.array([L.orc_det_cosf(float(y)) for y in ys], f32)
    np.testing.assert_array_equal(got, np.cos(ys.astype(np.float64)).astype(f32))
