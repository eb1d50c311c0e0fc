#!/usr/bin/env python3
"""One config-2 registration (after --warm warm-up registrations) for ncu captures:
   ncu --set full -k regex:icp_iteration_kernel -s <31*warm + launch> -c 1 python tools/one_registration.py --warm 1
--kind p2plane|gicp, --points N."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--kind", default="p2plane")
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    import cupoch_b200 as cph
    from cupoch_b200.testing import datagen
    R = cph.registration
    tgt, tn = datagen.surface(a.points, 11)
    src, sn = datagen.make_source(tgt, datagen.gt_transform(), 13, 14, 5e-4, attrs=[(tn, True)])
    s, t = cph.geometry.PointCloud(src), cph.geometry.PointCloud(tgt)
    t.normals = tn
    crit = R.ICPConvergenceCriteria(0, 0, a.iters)
    if a.kind == "gicp":
        s.normals = sn
        run = lambda: R.registration_generalized_icp(s, t, 0.02, np.eye(4), None, crit, return_correspondences=False)
    else:
        run = lambda: R.registration_icp(s, t, 0.02, np.eye(4), R.TransformationEstimationPointToPlane(), crit,
                                         return_correspondences=False)
    for _ in range(a.warm + 1):
        r = run()
    print("fitness %.6f rmse %.6g loop_ms %.3f launches %d" % (r.fitness, r.inlier_rmse, r.loop_ms, r.loop_launches))


if __name__ == "__main__":
    main()
