import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cupoch_b200 as cph
from cupoch_b200.testing import datagen
R, G = cph.registration, cph.geometry
for n in (1_000_000, 2_000_000, 3_000_000, 5_000_000):
    tgt, tn = datagen.surface(n, 11)
    src, sn = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4, attrs=[(tn, True)])
    t_pc = G.PointCloud(tgt); t_pc.normals = tn
    s_pc = G.PointCloud(src); s_pc.normals = sn
    for name, fn in (("p2plane", lambda it: R.registration_icp(s_pc, t_pc, 0.02, np.eye(4), R.TransformationEstimationPointToPlane(), R.ICPConvergenceCriteria(0, 0, it), return_correspondences=False)),
                     ("gicp", lambda it: R.registration_generalized_icp(s_pc, t_pc, 0.02, np.eye(4), None, R.ICPConvergenceCriteria(0, 0, it), return_correspondences=False))):
        for it in (0, 1, 3):
            r = fn(it)
            print(n, name, "iters", it, "fitness %.4f rmse %.6f" % (r.fitness, r.inlier_rmse), "T finite", bool(np.isfinite(r.transformation).all()), flush=True)
    sc = R._with_covariances(s_pc, 1e-3).covariances.cpu(); tc = R._with_covariances(t_pc, 1e-3).covariances.cpu()
    print(n, "cov finite", bool(np.isfinite(sc).all() and np.isfinite(tc).all()), "normals norm", float(np.abs(np.linalg.norm(tn, axis=1) - 1).max()), flush=True)
