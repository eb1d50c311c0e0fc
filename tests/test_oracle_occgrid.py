"""The OccupancyGrid restatement (oracle/oracle.c) against the reference's own known-answer tests
(src/tests/geometry/occupancygrid.cpp:30-97: Bounds, GetVoxel, Insert, SetFreeArea) and a brute-force model of Insert."""
import numpy as np
import pytest

from oracle import oracle_py as orc


def test_bounds_reference_kat():            # occupancygrid.cpp:30-43
    g = orc.OccupancyGrid()
    assert np.isclose(g.voxel_size, 0.05) and g.resolution == 512
    g.voxel_size = 5.0
    g.add_voxels([[0, 0, 0]])
    g.add_voxels([[511, 511, 511]])
    np.testing.assert_array_equal(g.get_min_bound(), np.full(3, -512 * 5 * 0.5, np.float32))
    np.testing.assert_array_equal(g.get_max_bound(), np.full(3, 512 * 5 * 0.5, np.float32))


def test_get_voxel_reference_kat():         # occupancygrid.cpp:45-67
    g = orc.OccupancyGrid(1.0, 64)
    h = 32
    g.add_voxels([[h + 1, h, h]], True)
    known, p = g.get_voxel([1.5, 0.0, 0.0])
    assert known and np.float32(p) == np.float32(g.prob_hit_log)
    g.add_voxels([[h + 1, h, h]], True)
    known, p = g.get_voxel([1.5, 0.0, 0.0])
    assert known and np.isclose(p, 2.0 * g.prob_hit_log)
    g.add_voxels([[h + 1, h, h]], False)
    known, p = g.get_voxel([1.5, 0.0, 0.0])
    assert known and np.isclose(p, 2.0 * g.prob_hit_log + g.prob_miss_log)
    assert g.get_voxel([1000.0, 0, 0]) == (False, pytest.approx(float("nan"), nan_ok=True))


def test_insert_reference_kat():            # occupancygrid.cpp:69-89
    g = orc.OccupancyGrid(1.0, 64, (-0.5, -0.5, 0))
    g.insert([[0.0, 0.0, 3.5]], [0, 0, 0])
    idx, prob = g.extract(0)
    assert len(idx) == 4
    for z, want in ((0.5, True), (1.5, True), (2.5, True), (3.5, True), (4.5, False)):
        assert g.get_voxel([0.0, 0.0, z])[0] == want
    assert np.float32(g.get_voxel([0, 0, 3.5])[1]) == np.float32(g.prob_hit_log)       # the end voxel is occupied
    assert np.float32(g.get_voxel([0, 0, 1.5])[1]) == np.float32(g.prob_miss_log)      # the ray's voxels are free
    assert len(g.extract(1)[0]) == 3 and len(g.extract(2)[0]) == 1


def test_set_free_area_reference_kat():     # occupancygrid.cpp:91-97
    g = orc.OccupancyGrid()
    g.set_free_area([0, 0, 0], [0.1, 0.1, 0.1])
    idx, prob = g.extract(1)
    assert len(idx) == 27
    assert (idx == 0).all()                 # SetFreeArea never writes grid_index_ (occupancygrid.cu:441-446)
    assert (prob == np.float32(g.prob_miss_log)).all()


def test_insert_each_voxel_once_and_max_range():
    """many rays through the same voxels: one update per voxel per Insert (sort + unique in the reference), occupied wins
    over free (set_difference), rays beyond max_range are cut and leave no occupied voxel"""
    rng = np.random.default_rng(5)
    g = orc.OccupancyGrid(0.1, 128)
    pts = (rng.random((3000, 3)).astype(np.float32) - 0.5) * 8.0
    g.insert(pts, [0.05, 0.05, 0.05], max_range=3.0)
    idx, prob = g.extract(0)
    assert set(np.unique(prob).tolist()) <= {np.float32(g.prob_hit_log), np.float32(g.prob_miss_log)}
    d = np.linalg.norm(pts - np.float32(0.05), axis=1)
    occ_idx, _ = g.extract(2)
    want = {tuple((np.floor(p / np.float32(0.1)).astype(int) + 64).tolist()) for p in pts[d <= 3.0]}
    assert {tuple(v) for v in occ_idx.tolist()} == want
    assert (g._stamp == 0).all()
    # a second identical insert adds exactly one more increment everywhere (clamped)
    g.insert(pts, [0.05, 0.05, 0.05], max_range=3.0)
    _, prob2 = g.extract(0)
    lo, hi = np.float32(g.clamping_thres_min), np.float32(g.clamping_thres_max)
    want2 = np.where(prob > 0, np.minimum(prob + np.float32(g.prob_hit_log), hi), np.maximum(prob + np.float32(g.prob_miss_log), lo))
    np.testing.assert_array_equal(prob2, want2.astype(np.float32))
