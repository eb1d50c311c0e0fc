"""GPU parity: KDTreeFlann search (C ABI through the Python mirror) vs the CPU oracle.
Indices and squared distances must be bit-exact (same arithmetic contract, same tie rule)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph
from cupoch_b200.testing import datagen


def _check(orc, tgt, qry, k, radius=None):
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    if radius is None:
        cnt, idx, d2 = tree.search_knn(qry, k)
        oi, od, oc = orc.search(tgt, qry, k, kdtree=True)
    else:
        cnt, idx, d2 = tree.search_radius(qry, radius, k)
        oi, od, oc = orc.search(tgt, qry, k, radius=radius, kdtree=True)
    idx, d2 = idx.cpu(), d2.cpu()
    assert cnt == oc
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(d2.view(np.uint32), od.view(np.uint32))
    return idx, d2


def test_golden_knn(golden, orc):
    g = golden["knn"]
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(np.array(g["points"], np.float32)))
    k, idx, d2 = tree.search_knn_vector_3f(g["query"], g["k"])
    assert k == g["result"]
    assert sorted(idx.tolist()) == sorted(g["indices"])
    np.testing.assert_allclose(np.sort(d2), np.sort(np.array(g["distance2"], np.float32)), atol=1e-4)


def test_golden_radius(golden, orc):
    g = golden["radius"]
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(np.array(g["points"], np.float32)))
    k, idx, d2 = tree.search_radius_vector_3f(g["query"], g["radius"], g["max_nn"])
    assert k == g["result"]
    assert sorted(idx.tolist()) == sorted(g["indices"])
    np.testing.assert_allclose(np.sort(d2), np.sort(np.array(g["distance2"], np.float32)), atol=1e-4)
    k2, idx2, _ = tree.search_vector_3f(g["query"], cph.geometry.KDTreeSearchParamRadius(g["radius"], g["max_nn"]))
    assert k2 == k and idx2.tolist() == idx.tolist()


@pytest.mark.parametrize("m,n", [(1, 1), (5, 40), (31, 33), (32, 32), (33, 31), (1000, 777), (20000, 30000)])
def test_radius_1nn_uniform(orc, m, n):
    tgt = datagen.uniform_cube(m, 100 + m)
    qry = datagen.uniform_cube(n, 200 + n, lo=(-0.1, -0.1, -0.1), hi=(1.1, 1.1, 1.1))
    for r in (0.02, 0.08, 2.5):
        _check(orc, tgt, qry, 1, radius=r)


def test_radius_1nn_surface_misaligned(orc):
    tgt, _ = datagen.surface(60000, 11)
    src = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4)
    idx, _ = _check(orc, tgt, src, 1, radius=0.02)
    assert (idx < 0).any() and (idx >= 0).any()


@pytest.mark.parametrize("k", [2, 7, 30, 100])
def test_knn_k(orc, k):
    tgt = datagen.uniform_cube(5000, 5)
    qry = datagen.uniform_cube(1500, 6)
    _check(orc, tgt, qry, k)
    _check(orc, tgt, qry, k, radius=0.07)


def test_knn_fewer_points_than_k(orc):
    tgt = datagen.uniform_cube(9, 7)
    qry = datagen.uniform_cube(50, 8)
    idx, d2 = _check(orc, tgt, qry, 20)
    assert (idx[:, 9:] == -1).all() and np.isinf(d2[:, 9:]).all()


def test_ties_and_duplicates(orc):
    # lattice target + duplicated points; queries at cell centres: many exact ties
    g = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3) * 0.25
    tgt = np.concatenate([g, g[::3], g[::7]]).astype(np.float32)
    rng = np.random.default_rng(3)
    tgt = tgt[rng.permutation(len(tgt))]
    qry = np.concatenate([g + 0.125, g, g + np.float32(0.125) * np.array([1, 0, 0], np.float32)]).astype(np.float32)
    _check(orc, tgt, qry, 1, radius=0.5)
    _check(orc, tgt, qry, 6, radius=0.3)
    _check(orc, tgt, qry, 9)


def test_strict_radius_boundary(orc):
    tgt = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0]], np.float32)
    qry = np.array([[0.5, 0, 0], [0, 1, 0]], np.float32)
    _check(orc, tgt, qry, 2, radius=0.5)    # d2 == r2 exactly -> excluded
    idx, _ = _check(orc, tgt, qry, 2, radius=1.0)
    assert idx[0].tolist() == [0, 1]           # tie at d2 = 0.25: smaller index first
    assert idx[1].tolist() == [-1, -1]         # both neighbours of (0,1,0) sit at d2 == r2 == 1.0: excluded
    # radius 0: r2 == 0, nothing satisfies d2 < 0 (the oracle wrapper reserves radius<=0 for "no radius")
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    cnt, idx, d2 = tree.search_radius(tgt, 0.0, 1)
    assert cnt == 0 and (idx.cpu() == -1).all() and np.isinf(d2.cpu()).all()


def test_error_codes():
    tree = cph.geometry.KDTreeFlann()
    assert tree.search_knn(np.zeros((3, 3), np.float32), 1)[0] == -1          # no data
    tree.set_geometry(cph.geometry.PointCloud(datagen.uniform_cube(10, 1)))
    assert tree.search_knn(np.zeros((0, 3), np.float32), 1)[0] == -1          # empty query
    assert tree.search_knn(np.zeros((2, 3), np.float32), 101)[0] == -1        # k > NUM_MAX_NN
    with pytest.raises(RuntimeError):
        tree.search_knn_vector_3f([0, 0, 0], -1)


def test_self_query_1m_property():
    # full-size property: every point of a 1M cloud finds itself at d2 == 0 (points are distinct)
    tgt = datagen.uniform_cube(1_000_000, 21)
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    cnt, idx, d2 = tree.search_radius(tgt, 0.05, 1)
    idx, d2 = idx.cpu()[:, 0], d2.cpu()[:, 0]
    assert cnt == len(tgt)
    same = idx == np.arange(len(tgt))
    # duplicates (if any) resolve to the smaller index
    assert (d2 == 0).all() and ((idx <= np.arange(len(tgt))).all())
    assert same.mean() > 0.999
