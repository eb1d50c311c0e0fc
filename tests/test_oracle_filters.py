"""CPU oracle of the outlier filters (down_sample.cu:317-438) against the reference's known-answer test
(tests/geometry/pointcloud.cpp:676-693, via tests/golden) and against an independent float64 numpy restatement."""
import numpy as np


def test_radius_outliers_golden(golden, orc):
    g = golden["radius_outliers"]
    pts = np.array(g["points"], np.float32)
    kept = orc.remove_radius_outliers(pts, g["nb_points"], g["radius"])
    assert kept.tolist() == [0]
    np.testing.assert_array_equal(pts[kept], np.array(g["kept_points"], np.float32))


def _brute_d2(pts):
    p = pts.astype(np.float64)
    return ((p[:, None, :] - p[None]) ** 2).sum(-1)


def test_radius_outliers_vs_numpy(orc):
    rng = np.random.default_rng(3)
    pts = rng.random((3000, 3), dtype=np.float32)
    pts[:40] += 2.0  # isolated points
    d2 = _brute_d2(pts)
    for nb, r in ((5, 0.06), (16, 0.1), (0, 0.01), (1, 0.05)):
        kept = orc.remove_radius_outliers(pts, nb, r)
        rr = np.float32(r) * np.float32(r)
        ref = np.flatnonzero((d2 < float(rr) * (1 - 1e-6)).sum(1) > nb)
        amb = np.flatnonzero(((d2 < float(rr) * (1 + 1e-6)).sum(1) > nb) != ((d2 < float(rr) * (1 - 1e-6)).sum(1) > nb))
        assert set(kept) - set(amb) == set(ref) - set(amb)
        assert (np.diff(kept) > 0).all()
    assert len(orc.remove_radius_outliers(pts, 5, 0.0)) == 0                     # radius 0 matches nothing
    np.testing.assert_array_equal(orc.remove_radius_outliers(pts, 5, -0.06),     # the search squares the radius
                                  orc.remove_radius_outliers(pts, 5, 0.06))


def test_statistical_outliers_vs_numpy(orc):
    rng = np.random.default_rng(4)
    pts = rng.random((2500, 3), dtype=np.float32)
    pts[:30] += rng.random((30, 3), dtype=np.float32) * 3 + 1.5
    d2 = np.sort(_brute_d2(pts), 1)
    for k, ratio in ((16, 1.0), (8, 2.0), (1, 1.0)):
        kept, avg, (mean, std, thr) = orc.remove_statistical_outliers(pts, k, ratio)
        ref_avg = d2[:, :k].mean(1)     # squared distances, the point itself (0) included -- as the reference
        np.testing.assert_allclose(avg, ref_avg, rtol=2e-5, atol=1e-9)
        ok = ref_avg > 0
        m = ref_avg[ref_avg >= 0].mean()
        s = np.sqrt((((ref_avg - m) ** 2) * ok).sum() / (len(ref_avg) - 1))
        np.testing.assert_allclose([mean, std, thr], [m, s, m + ratio * s], rtol=1e-4)
        margin = 1e-4 * thr
        sure_in = np.flatnonzero((ref_avg > 0) & (ref_avg < thr - margin))
        sure_out = np.flatnonzero(~((ref_avg > 0) & (ref_avg < thr + margin)))
        assert set(sure_in) <= set(kept) and not (set(sure_out) & set(kept))
        assert (np.diff(kept) > 0).all()
        if k == 1:
            assert len(kept) == 0       # only the point itself: every mean is 0 and `dist > 0` drops it
    assert len(pts) - len(orc.remove_statistical_outliers(pts, 16, 1.0)[0]) >= 25  # the planted outliers go


def test_voxel_grid_oracle(orc):
    """voxelgrid_factory.cu:164-228 vs numpy; tests/geometry/voxelgrid.cpp:57-68 (one point inside wide bounds ->
    one voxel) as the known answer."""
    k, c, o = orc.voxel_grid_from_point_cloud(np.array([[0.5, 0.5, 0.5]], np.float32), 1.0, [-100] * 3, [100] * 3)
    assert len(k) == 1 and k.tolist() == [[100, 100, 100]] and c.tolist() == [[1.0, 1.0, 1.0]]
    rng = np.random.default_rng(8)
    pts = (rng.random((5000, 3), dtype=np.float32) * np.float32(2) - np.float32(0.7)).astype(np.float32)
    col = rng.random((5000, 3), dtype=np.float32)
    for voxel, lo, hi in ((0.1, None, None), (0.07, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])):
        k, c, o = orc.voxel_grid_from_point_cloud(pts, voxel, lo, hi, colors=col)
        ki = np.floor((pts - o) / np.float32(voxel)).astype(np.int32)
        u, inv, cnt = np.unique(ki, axis=0, return_inverse=True, return_counts=True)
        np.testing.assert_array_equal(k, u)                       # lexicographic, negative indices included
        ref = np.zeros((len(u), 3))
        np.add.at(ref, inv.reshape(-1), col.astype(np.float64))
        np.testing.assert_allclose(c, ref / cnt[:, None], rtol=1e-6)
        if lo is not None:
            assert k.min() < 0                                    # points below the bound are not clipped
    assert len(orc.voxel_grid_from_point_cloud(pts, 0.0)[0]) == 0


def test_gaussian_filter_oracle(orc):
    """pointcloud.cu:56-106,387-433 vs a float64 numpy restatement"""
    rng = np.random.default_rng(6)
    pts = rng.random((2000, 3), dtype=np.float32)
    col = rng.random((2000, 3), dtype=np.float32)
    r, sigma2, k = 0.12, 0.003, 40
    op, on, oc = orc.gaussian_filter(pts, r, sigma2, k, colors=col)
    assert on is None and op.shape == pts.shape
    d2 = _brute_d2(pts)
    rr = float(np.float32(r) * np.float32(r))
    for i in range(0, 2000, 37):
        order = np.argsort(d2[i], kind="stable")[:k]
        order = order[d2[i][order] < rr]
        w = np.exp(-0.5 * d2[i][order] / float(np.float32(sigma2)))
        np.testing.assert_allclose(op[i], (w[:, None] * pts[order]).sum(0) / w.sum(), rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(oc[i], (w[:, None] * col[order]).sum(0) / w.sum(), rtol=2e-5, atol=1e-6)
    for bad in ((0.0, sigma2, k), (r, 0.0, k), (r, sigma2, 0)):
        assert len(orc.gaussian_filter(pts, *bad)[0]) == 0
