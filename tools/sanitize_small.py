#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cupoch_b200 as cph
from cupoch_b200.testing import datagen

R, G = cph.registration, cph.geometry
n = int(os.environ.get('SANITIZE_POINTS', '20000'))   # >= 8192 points: the certified regime runs with helper blocks
tgt, tn = datagen.surface(n, 11)
tc = datagen.texture(tgt)
src, sn, sc = datagen.make_source(tgt, datagen.gt_transform((-1, 1.5, 2), (0.01, -0.005, 0.008)), 13, 14, 5e-4,
                                  attrs=[(tn, True), (tc, False)])
s, t = G.PointCloud(src), G.PointCloud(tgt)
s.normals, s.colors, t.normals, t.colors = sn, sc, tn, tc
crit = R.ICPConvergenceCriteria(0, 0, 12)         # 13 launches: searching regime, then the certified one (DMMA, cp.async pipeline, fused tail)
for name, fn in (("p2p", lambda: R.registration_icp(s, t, 0.03, np.eye(4), R.TransformationEstimationPointToPoint(), crit)),
                 ("p2plane", lambda: R.registration_icp(s, t, 0.03, np.eye(4), R.TransformationEstimationPointToPlane(), crit)),
                 ("symmetric", lambda: R.registration_icp(s, t, 0.03, np.eye(4), R.TransformationEstimationSymmetricMethod(), crit)),
                 ("gicp", lambda: R.registration_generalized_icp(s, t, 0.03, np.eye(4), None, crit)),
                 ("colored", lambda: R.registration_colored_icp(s, t, 0.03, np.eye(4), crit))):
    r = fn()
    print(name, "fitness %.3f" % r.fitness, "corr", len(r.correspondence_set))
tree = G.KDTreeFlann(t)
print("knn", tree.search_knn(s.points, 8)[0], "radius", tree.search_radius(s.points, 0.01, 5)[0])
d = t.voxel_down_sample(0.05)
print("voxel", len(d))
t2 = G.PointCloud(tgt)
t2.estimate_normals(G.KDTreeSearchParamKNN(10))
print("normals ok", bool(np.isfinite(t2.normals.cpu()).all()))
t2.estimate_normals(G.KDTreeSearchParamRadius(0.03, 20))
print("normals (radius) ok", bool(np.isfinite(t2.normals.cpu()).all()))
og = G.OccupancyGrid(0.05, 96)
rng = np.random.default_rng(3)
scan = (rng.standard_normal((5000, 3)).astype(np.float32) * 1.2).astype(np.float32)
og.insert(scan, (0.1, 0.0, -0.1), 2.0)
og.add_voxels(rng.integers(0, 96, (500, 3)).astype(np.int32), True)
og.set_free_area((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5))
print("occupancy grid known voxels", len(og.extract_known_voxels()[1]), "occupied", len(og.extract_occupied_voxels()[1]))
small = G.PointCloud(tgt[:4000])
small.normals = tn[:4000]
print("fpfh", R.compute_fpfh_feature(small, G.KDTreeSearchParamKNN(10)).cpu().shape, "dbscan clusters", int(small.cluster_dbscan(0.02, 5).cpu().max()) + 1)
out, idx = small.remove_radius_outlier(4, 0.02)
print("radius outlier kept", len(out), "gaussian", len(small.gaussian_filter(0.02, 1e-4, 20)))
