"""GPU parity: PointCloud ops (VoxelDownSample, EstimateNormals, Transform, bounds) and the GICP /
Colored-ICP initialisers vs the reference's golden vectors and the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cupoch_b200 as cph
from conftest import lexsort_rows
from cupoch_b200.testing import datagen

TOL = 1e-4


def test_voxel_golden(golden):
    g = golden["voxel"]
    pc = cph.geometry.PointCloud(np.array(g["points"], np.float32))
    pc.normals = np.array(g["normals"], np.float32)
    pc.colors = np.array(g["colors"], np.float32)
    out = pc.voxel_down_sample(g["voxel_size"])
    assert len(out) == 20
    np.testing.assert_allclose(lexsort_rows(out.points.cpu()), lexsort_rows(g["ref_points"]), atol=TOL, rtol=0)
    np.testing.assert_allclose(lexsort_rows(out.normals.cpu()), lexsort_rows(g["ref_normals"]), atol=TOL, rtol=0)
    np.testing.assert_allclose(lexsort_rows(out.colors.cpu()), lexsort_rows(g["ref_colors"]), atol=TOL, rtol=0)


# voxel 0.05: 16 k cells for 200 k points -> the dense-grid path; 0.004: 31 M cells -> the hash path; "forced": the hash
# path on the dense case (CPHB_VOXEL_NO_DENSE), so both paths see the same inputs
@pytest.mark.parametrize("voxel,force_hash", [(0.05, False), (0.05, True), (0.004, False)])
@pytest.mark.parametrize("attrs", ["p", "pn", "pc", "pnc"])
def test_voxel_vs_oracle(orc, attrs, voxel, force_hash, monkeypatch):
    n = 200000
    p = datagen.uniform_cube(n, 21, hi=(2, 2, 0.5))
    nr = datagen.unit_normals(n, 22) if "n" in attrs else None
    co = datagen.uniform_cube(n, 23) if "c" in attrs else None
    pc = cph.geometry.PointCloud(p)
    pc.normals, pc.colors = nr, co
    if force_hash:
        monkeypatch.setenv("CPHB_VOXEL_NO_DENSE", "1")
    out = pc.voxel_down_sample(voxel)
    op, on, oc = orc.voxel_down_sample(p, voxel, nr, co)
    assert len(out) == len(op)
    # same voxel order (lexicographic) and float64-accumulated means: bit-exact
    np.testing.assert_array_equal(out.points.cpu(), op)
    if nr is not None:
        np.testing.assert_array_equal(out.normals.cpu(), on)
    if co is not None:
        np.testing.assert_array_equal(out.colors.cpu(), oc)


def test_voxel_rejects_and_edges(orc):
    p = datagen.uniform_cube(1000, 3)
    pc = cph.geometry.PointCloud(p)
    assert len(pc.voxel_down_sample(0.0)) == 0
    assert len(pc.voxel_down_sample(-1.0)) == 0
    assert len(pc.voxel_down_sample(1e-12)) == 0
    one = pc.voxel_down_sample(10.0)
    assert len(one) == 1
    np.testing.assert_allclose(one.points.cpu()[0], p.astype(np.float64).mean(0), rtol=1e-6)
    # idempotence at full coverage: down-sampling the voxel centres again keeps the count
    d1 = pc.voxel_down_sample(0.1)
    assert len(cph.geometry.PointCloud(d1.points.cpu()).voxel_down_sample(1e-4)) == len(d1)


def test_bounds_golden(golden):
    g = golden["bounds"]
    pc = cph.geometry.PointCloud(np.array(g["points"], np.float32))
    np.testing.assert_allclose(pc.get_min_bound(), g["min"], atol=TOL)
    np.testing.assert_allclose(pc.get_max_bound(), g["max"], atol=TOL)


def test_transform_golden_and_oracle(golden, orc):
    g = golden["transform"]
    p = np.array(g["points"], np.float32)
    c, s = np.cos(np.pi / 4), np.sin(np.pi / 4)
    T = np.array([[1, 0, 0, 1], [0, c, -s, 2], [0, s, c, 3], [0, 0, 0, 1]], np.float32)
    Ti = np.eye(4, dtype=np.float32)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -(Ti[:3, :3] @ T[:3, 3])
    pc = cph.geometry.PointCloud(p)
    pc.normals = p
    pc.transform(T)
    np.testing.assert_array_equal(pc.points.cpu(), orc.transform_points(p, T))       # bit-exact vs oracle
    np.testing.assert_array_equal(pc.normals.cpu(), orc.transform_normals(p, T))
    pc.transform(Ti)
    np.testing.assert_allclose(pc.points.cpu(), p, atol=g["tolerance"], rtol=0)
    np.testing.assert_allclose(pc.normals.cpu(), p, atol=g["tolerance"], rtol=0)
    cov = orc.covariances_from_normals(datagen.unit_normals(100, 1))
    pc2 = cph.geometry.PointCloud(datagen.uniform_cube(100, 2))
    pc2.covariances = cov
    pc2.transform(T)
    np.testing.assert_array_equal(pc2.covariances.cpu(), orc.rotate_covariances(cov, T))


def test_estimate_normals_golden(golden):
    g = golden["normals"]
    pc = cph.geometry.PointCloud(np.array(g["points"], np.float32))
    pc.estimate_normals(cph.geometry.KDTreeSearchParamKNN(g["knn"]))
    n = pc.normals.cpu()
    ref = np.array(g["ref"], np.float32)
    flip = np.sign(ref[:, 0]) * np.sign(n[:, 0]) < 0
    n[flip] *= -1
    np.testing.assert_allclose(n, ref, atol=TOL, rtol=0)


def test_estimate_normals_vs_oracle(orc):
    p, nt = datagen.surface(30000, 5)
    pc = cph.geometry.PointCloud(p)
    pc.estimate_normals(cph.geometry.KDTreeSearchParamKNN(20))
    n = pc.normals.cpu()
    o = orc.estimate_normals(p, knn=20)
    # same neighbour lists, same cumulant order, shared deterministic acos / cos inside FastEigen3x3: bit-exact
    np.testing.assert_array_equal(n, o)
    assert (np.abs((n * nt).sum(1)) > 0.95).mean() > 0.98         # and they are normals of the surface
    pc.estimate_normals(cph.geometry.KDTreeSearchParamRadius(0.02, 30))
    o = orc.estimate_normals(p, knn=0, radius=0.02, max_nn=30)
    n = pc.normals.cpu()
    np.testing.assert_array_equal(n, o)


def test_estimate_normals_fused_equals_two_pass(monkeypatch):
    """the fused kernel (search + cumulants + eigen-solve, no [n][k] table) against the two-pass form it replaces
    (CPHB_NORMALS_UNFUSED=1): bit for bit, kNN and radius, including points with < 3 neighbours, a cloud that does not
    fill its last leaf, and the block form used by the sharded path"""
    import ctypes as C
    from cupoch_b200 import _lib
    from cupoch_b200.utility import DeviceArray
    p = np.concatenate([datagen.surface(50_017, 6)[0], datagen.uniform_cube(300, 7) * 3 + 2]).astype(np.float32)  # + isolated points
    for param in (cph.geometry.KDTreeSearchParamKNN(30), cph.geometry.KDTreeSearchParamRadius(0.015, 30),
                  cph.geometry.KDTreeSearchParamKNN(2)):
        pc = cph.geometry.PointCloud(p)
        pc.estimate_normals(param)
        fused = pc.normals.cpu()
        monkeypatch.setenv("CPHB_NORMALS_UNFUSED", "1")
        pc.estimate_normals(param)
        two = pc.normals.cpu()
        monkeypatch.delenv("CPHB_NORMALS_UNFUSED")
        np.testing.assert_array_equal(fused, two)
    assert (fused[-300:] == [0, 0, 1]).all(1).any()          # knn=2 < 3 neighbours: the (0,0,1) branch ran
    # block form (queries in the block's Hilbert order instead of index order)
    d = DeviceArray.from_numpy(p)
    out = DeviceArray((20_000, 3), np.float32)
    _lib.check(_lib.lib().cphb_estimate_normals_range(d.ptr, len(p), 30, 0.0, 0, 10_000, 20_000, out.ptr, None))
    pc = cph.geometry.PointCloud(p)
    pc.estimate_normals(cph.geometry.KDTreeSearchParamKNN(30))
    np.testing.assert_array_equal(out.cpu(), pc.normals.cpu()[10_000:30_000])


def test_gicp_covariances_bit_exact(orc):
    nrm = datagen.unit_normals(5000, 9)
    nrm[0] = [-1, 0, 0]
    nrm[1] = [1, 0, 0]
    pc = cph.geometry.PointCloud(datagen.uniform_cube(5000, 1))
    pc.normals = nrm
    out = cph.registration._with_covariances(pc, 1e-3)
    np.testing.assert_array_equal(out.covariances.cpu(), orc.covariances_from_normals(nrm, 1e-3))


def test_color_gradient_bit_exact(orc):
    p, nt = datagen.surface(20000, 31)
    col = datagen.texture(p)
    pc = cph.geometry.PointCloud(p)
    pc.normals, pc.colors = nt, col
    out = cph.registration.initialize_pointcloud_for_colored_icp(pc, 0.04, 30)
    nbr, _, _ = orc.search(p, p, 30, radius=0.04, kdtree=True)
    np.testing.assert_array_equal(out._color_gradient.cpu(), orc.color_gradient(p, nt, col, nbr))
