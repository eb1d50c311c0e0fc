"""GPU parity against the REFERENCE'S OWN search engine: the FLANN CUDA kd-tree cupoch vendors
(third_party/flann/algorithms/kdtree_cuda_3d_index.cu), compiled from where it lies under /root/reference into
oracle/_ref/libflann_ref.so (oracle/ref_flann/Makefile; `__graft_entry__.build()` builds it when the reference is
present, the GPU box uses the prebuilt file) and called exactly like cupoch::knn::KDTreeFlann calls it
(kdtree_flann.inl:70-144).  This pins the search row to reference CODE, not to a restatement:

  * which point is returned (index) and its d2 -- BIT-IDENTICAL: the product computes d2 in the operation order the
    reference's kernel has in SASS (FMUL dy,dy; FFMA dx,dx; FFMA dz,dz) -- on random clouds and on the ICP workload;
  * the tie rule (DESIGN.md hazard 1): FLANN keeps the first-visited of equal-distance points in kd order, the
    product keeps the smallest index -- measured here, asserted only through d2 (the distances must agree);
  * k > 1 radius results that are not full (R1b, result_set.h:405-473): same SET of neighbours.

Statistics of every comparison are written to gpurun_out/flann_parity.json (copied to profiles/ by hand).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph
from cupoch_b200.testing import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libflann_ref.so")
STATS = {}


@pytest.fixture(scope="module")
def fref():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libflann_ref.so not built (needs /root/reference at build time)")
    L = C.CDLL(SO)
    L.fref_build.restype = C.c_void_p
    L.fref_build.argtypes = [C.c_void_p, C.c_int]
    L.fref_free.argtypes = [C.c_void_p]
    L.fref_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.fref_radius.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    yield L
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "flann_parity.json"), "w") as f:
        json.dump(STATS, f, indent=1)


class Flann:
    def __init__(self, L, tgt):
        self.L = L
        self.tgt = np.ascontiguousarray(tgt, np.float32)
        self.h = L.fref_build(self.tgt.ctypes.data, len(self.tgt))
        assert self.h, "the reference's FLANN index failed to build"

    def radius(self, q, r, k):
        q = np.ascontiguousarray(q, np.float32)
        idx = np.empty((len(q), k), np.int32)
        d2 = np.empty((len(q), k), np.float32)
        assert self.L.fref_radius(self.h, q.ctypes.data, len(q), C.c_float(r), k, idx.ctypes.data, d2.ctypes.data) == 0
        return idx, d2

    def knn(self, q, k):
        q = np.ascontiguousarray(q, np.float32)
        idx = np.empty((len(q), k), np.int32)
        d2 = np.empty((len(q), k), np.float32)
        assert self.L.fref_knn(self.h, q.ctypes.data, len(q), k, idx.ctypes.data, d2.ctypes.data) == 0
        return idx, d2

    def close(self):
        self.L.fref_free(self.h)


def _ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def _compare_1nn(name, fref, tgt, qry, r):
    ref = Flann(fref, tgt)
    fi, fd = ref.radius(qry, r, 1)
    ref.close()
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    _, idx, d2 = tree.search_radius(qry, r, 1)
    idx, d2 = idx.cpu()[:, 0], d2.cpu()[:, 0]
    fi, fd = fi[:, 0], fd[:, 0]
    found, ffound = idx >= 0, fi >= 0
    both = found & ffound
    # a query whose nearest point lies within an ulp of the radius may be in on one side and out on the other
    # (FLANN's d2 is compiled with --use_fast_math contraction, ours follows the written contract)
    edge = found != ffound
    u = _ulps(d2[both], fd[both])
    idx_diff = both & (idx != fi)
    # where the indices differ the two points must be (numerically) equidistant: a tie, not a wrong answer
    tie_ok = True
    if idx_diff.any():
        q = qry[idx_diff].astype(np.float64)
        da = ((q - tgt[idx[idx_diff]].astype(np.float64)) ** 2).sum(1)
        db = ((q - tgt[fi[idx_diff]].astype(np.float64)) ** 2).sum(1)
        tie_ok = bool(np.all(np.abs(da - db) <= 4e-7 * np.maximum(da, db) + 1e-30))
    STATS[name] = {"queries": int(len(qry)), "targets": int(len(tgt)), "radius": float(r), "found_product": int(found.sum()),
                   "found_flann": int(ffound.sum()), "membership_differs": int(edge.sum()),
                   "index_differs": int(idx_diff.sum()), "index_differs_all_ties": tie_ok,
                   "d2_bit_equal": int((u == 0).sum()), "d2_max_ulps": int(u.max()) if len(u) else 0}
    assert tie_ok, "product and FLANN return different, non-equidistant points"
    assert edge.sum() <= max(2, len(qry) // 100000)
    if edge.any():  # those must sit on the radius
        de = np.where(found[edge], d2[edge], fd[edge])
        assert np.all(np.abs(de - np.float32(r) * np.float32(r)) <= 4e-7 * de)
    # d2 follows the operation order the reference's kernel has in SASS (cphb_internal.cuh dist2): bit-identical
    assert (u == 0).all(), "d2 differs from the reference binary's in %d of %d results" % (int((u != 0).sum()), len(u))
    return STATS[name]


def test_flann_1nn_uniform(fref):
    tgt = datagen.uniform_cube(200_000, 101)
    qry = datagen.uniform_cube(100_000, 202, lo=(-0.05, -0.05, -0.05), hi=(1.05, 1.05, 1.05))
    s = _compare_1nn("uniform_200k_r0.02", fref, tgt, qry, 0.02)
    assert s["index_differs"] <= 5          # random floats: ties are essentially impossible
    _compare_1nn("uniform_200k_r0.3", fref, tgt, qry, 0.3)


def test_flann_1nn_icp_workload(fref):
    # the first search of config 2 (misaligned surface), 1 M -> 1 M
    tgt, _ = datagen.surface(1_000_000, 11)
    src = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4)
    s = _compare_1nn("config2_first_search_1M", fref, tgt, src, 0.02)
    assert s["index_differs"] <= 20


def test_flann_ties(fref):
    # lattice + duplicates, queries at cell centres / on lattice points: every query has several equidistant
    # nearest points.  d2 must agree bit for bit (all values are exact in float); the INDEX is the tie rule.
    g = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3) * 0.25
    tgt = np.concatenate([g, g[::3], g[::7]]).astype(np.float32)
    tgt = tgt[np.random.default_rng(3).permutation(len(tgt))]
    qry = np.concatenate([g + 0.125, g]).astype(np.float32)
    ref = Flann(fref, tgt)
    fi, fd = ref.radius(qry, 0.5, 1)
    ref.close()
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    _, idx, d2 = tree.search_radius(qry, 0.5, 1)
    idx, d2 = idx.cpu(), d2.cpu()
    np.testing.assert_array_equal(d2, fd)
    # both answers are nearest points; ours is the smallest index among them
    d_ours = ((qry - tgt[idx[:, 0]]) ** 2).sum(1)
    d_ref = ((qry - tgt[fi[:, 0]]) ** 2).sum(1)
    np.testing.assert_array_equal(d_ours, d_ref)
    assert (idx[:, 0] <= fi[:, 0]).all()
    STATS["ties_lattice"] = {"queries": int(len(qry)), "same_index": int((idx == fi).sum()),
                             "product_index_smaller": int((idx < fi).sum())}


@pytest.mark.parametrize("k,r", [(4, 0.05), (15, 0.08), (30, 0.1)])
def test_flann_radius_k_not_full(fref, k, r):
    # R1b: radius search with max_nn > 1 where most result lists are NOT full.  The reference heap-sorts an array it
    # only heapifies on the k-th insert (result_set.h:405-473): compare as sets, then record whether its order is ascending
    tgt = datagen.uniform_cube(20_000, 5)
    qry = datagen.uniform_cube(5_000, 6)
    ref = Flann(fref, tgt)
    fi, fd = ref.radius(qry, r, k)
    ref.close()
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    _, idx, d2 = tree.search_radius(qry, r, k)
    idx, d2 = idx.cpu(), d2.cpu()
    same_set = np.array([set(a[a >= 0].tolist()) == set(b[b >= 0].tolist()) for a, b in zip(idx, fi)])
    n_found = (idx >= 0).sum(1)
    full = n_found == k
    ref_sorted = np.array([bool(np.all(np.diff(d[i >= 0]) >= 0)) for i, d in zip(fi, fd)])
    ref_prefix = np.array([bool(np.all(i[:c] >= 0)) for i, c in zip(fi, (fi >= 0).sum(1))])
    STATS["radius_k%d_r%g" % (k, r)] = {"queries": int(len(qry)), "lists_full": int(full.sum()), "same_set": int(same_set.sum()),
                                        "ref_lists_ascending": int(ref_sorted.sum()), "ref_valid_entries_first": int(ref_prefix.sum()),
                                        "same_order": int(sum(np.array_equal(a, b) for a, b in zip(idx, fi)))}
    # lists that are not full contain EVERY point inside the radius on both sides: the sets must be equal.  A full
    # list may differ only where the k-th and (k+1)-th neighbour tie.
    assert same_set[~full].all()
    assert same_set.mean() > 0.999


def test_flann_knn30(fref):
    tgt = datagen.uniform_cube(50_000, 7)
    qry = datagen.uniform_cube(5_000, 8)
    ref = Flann(fref, tgt)
    fi, fd = ref.knn(qry, 30)
    ref.close()
    tree = cph.geometry.KDTreeFlann(cph.geometry.PointCloud(tgt))
    _, idx, d2 = tree.search_knn(qry, 30)
    idx, d2 = idx.cpu(), d2.cpu()
    same_set = np.array([set(a.tolist()) == set(b.tolist()) for a, b in zip(idx, fi)])
    STATS["knn30"] = {"queries": int(len(qry)), "same_set": int(same_set.sum()),
                      "same_order": int(sum(np.array_equal(a, b) for a, b in zip(idx, fi))),
                      "d2_max_ulps": int(_ulps(np.sort(d2, 1), np.sort(fd, 1)).max())}
    assert same_set.mean() > 0.999
    assert _ulps(np.sort(d2, 1), np.sort(fd, 1)).max() == 0
