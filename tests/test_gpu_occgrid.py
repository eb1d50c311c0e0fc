"""GPU parity of geometry::OccupancyGrid: the reference's own known-answer tests
(src/tests/geometry/occupancygrid.cpp:30-97) through the Python mirror / C ABI, and Insert / AddVoxels / SetFreeArea /
Extract* against the CPU oracle bit for bit (log-odds, grid indices, order, bounds) on random scans."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph

G = cph.geometry


def test_reference_kat_bounds():
    g = G.OccupancyGrid()
    assert np.isclose(g.voxel_size, 0.05) and g.resolution == 512
    g.origin = (0, 0, 0)
    g.voxel_size = 5
    g.add_voxel((0, 0, 0))
    g.add_voxel((511, 511, 511))
    np.testing.assert_array_equal(g.get_min_bound(), np.full(3, -512 * 5 * 0.5, np.float32))
    np.testing.assert_array_equal(g.get_max_bound(), np.full(3, 512 * 5 * 0.5, np.float32))
    with pytest.raises(Exception):
        g.add_voxel((512, 0, 0))


def test_reference_kat_get_voxel():
    g = G.OccupancyGrid(1.0, 512)
    h = 256
    for k, want in ((1, 0.85), (2, 1.7), (None, 1.7 - 0.4)):
        g.add_voxel((h + 1, h, h), k is not None)
        known, v = g.get_voxel((1.5, 0.0, 0.0))
        assert known and np.isclose(v.prob_log, want)
        assert v.grid_index.tolist() == [h + 1, h, h]
    assert not g.get_voxel((1e6, 0, 0))[0] and g.is_unknown((1e6, 0, 0)) and g.is_occupied((1.5, 0, 0))


def test_reference_kat_insert():
    g = G.OccupancyGrid(1.0, 512, (-0.5, -0.5, 0))
    g.insert(np.array([[0.0, 0.0, 3.5]], np.float32), (0, 0, 0))
    assert len(g.extract_known_voxels()[1]) == 4
    for z, want in ((0.5, True), (1.5, True), (2.5, True), (3.5, True), (4.5, False)):
        assert g.get_voxel((0.0, 0.0, z))[0] == want
    assert "with 4 voxels" in repr(g)


def test_reference_kat_set_free_area():
    g = G.OccupancyGrid()
    g.set_free_area((0, 0, 0), (0.1, 0.1, 0.1))
    idx, prob = g.extract_free_voxels()
    assert len(idx) == 27


def _same(g, o):
    for which, name in ((0, "known"), (1, "free"), (2, "occupied")):
        gi, gp = g._extract(which)
        oi, op = o.extract(which)
        np.testing.assert_array_equal(gi, oi, err_msg=name)
        np.testing.assert_array_equal(gp.view(np.uint32), op.view(np.uint32), err_msg=name)
    lo, hi = g._bounds()
    np.testing.assert_array_equal(np.concatenate([lo, hi]), o.bounds.astype(np.int32))
    np.testing.assert_array_equal(g.get_min_bound(), o.get_min_bound())
    np.testing.assert_array_equal(g.get_max_bound(), o.get_max_bound())


@pytest.mark.parametrize("max_range", [-1.0, 2.5])
def test_insert_vs_oracle(orc, max_range):
    rng = np.random.default_rng(11)
    res, vs, org = 160, 0.05, (0.1, -0.2, 0.05)
    g, o = G.OccupancyGrid(vs, res, org), orc.OccupancyGrid(vs, res, org)
    for k in range(3):                                       # three scans from different viewpoints accumulate and clamp
        d = rng.standard_normal((20000, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        vp = np.float32([0.3 * k, -0.1 * k, 0.2])
        pts = (vp + d * rng.uniform(0.5, 4.5, (len(d), 1)).astype(np.float32)).astype(np.float32)   # some leave the grid (+-4 m)
        pts[:5] = vp                                         # zero-length rays
        g.insert(G.PointCloud(pts), vp, max_range)
        o.insert(pts, vp, max_range)
        _same(g, o)
    # the whole dense array, not just the bound box
    np.testing.assert_array_equal(g.to_torch().cpu().numpy().reshape(-1).view(np.uint32), o.prob.view(np.uint32))


def test_add_voxels_set_free_area_clear_vs_oracle(orc):
    rng = np.random.default_rng(12)
    res = 96
    g, o = G.OccupancyGrid(0.1, res), orc.OccupancyGrid(0.1, res)
    v = rng.integers(0, res, (5000, 3)).astype(np.int32)
    v[:50] = v[0]                                             # the same voxel 50 times: 50 increments, clamped
    for occ in (True, False, True, True):
        g.add_voxels(v, occ)
        o.add_voxels(v, occ)
    _same(g, o)
    g.set_free_area((-1.0, -0.5, 0.0), (0.7, 0.9, 20.0))
    o.set_free_area((-1.0, -0.5, 0.0), (0.7, 0.9, 20.0))
    _same(g, o)
    g.prob_hit_log, o.prob_hit_log = 1.25, 1.25               # parameters are plain attributes
    g.add_voxels(v[:100], True)
    o.add_voxels(v[:100], True)
    _same(g, o)
    g.clear()
    assert len(g.extract_known_voxels()[1]) == 0
    lo, hi = g._bounds()
    assert (lo == res // 2).all() and (hi == res // 2).all()


def test_insert_large_scan_properties():
    """1 M rays into the default 512^3 grid: the size the reference's dense grid has; properties instead of an oracle run"""
    rng = np.random.default_rng(13)
    g = G.OccupancyGrid(0.05, 512)
    d = rng.standard_normal((1_000_000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(2.0, 10.0, (len(d), 1)).astype(np.float32)).astype(np.float32)
    g.insert(pts, (0, 0, 0))
    idx, prob = g.extract_known_voxels()
    assert set(np.unique(prob).tolist()) == {np.float32(g.prob_hit_log), np.float32(g.prob_miss_log)}
    occ, _ = g.extract_occupied_voxels()
    want = np.unique(np.floor(pts / np.float32(0.05)).astype(np.int64) + 256, axis=0)
    np.testing.assert_array_equal(np.unique(occ, axis=0), want)
    g.insert(pts, (0, 0, 0))                                  # idempotent up to one more (clamped) increment
    _, prob2 = g.extract_known_voxels()
    assert len(prob2) == len(prob)
    assert set(np.unique(prob2).tolist()) == {np.float32(np.float32(g.prob_hit_log) * 2), np.float32(np.float32(g.prob_miss_log) * 2)}
