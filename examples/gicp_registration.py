"""Counterpart of the reference's examples/python/basic/gicp_registration.py: Generalized ICP (covariances are derived from
estimated normals when the clouds have none, generalized_icp.cu:37-61)."""
import time

import numpy as np

from _clouds import pair
import cupoch_b200 as cph

if __name__ == "__main__":
    src, tgt, _, _, gt = pair()
    source_gpu = cph.geometry.PointCloud(src)
    target_gpu = cph.geometry.PointCloud(tgt)
    threshold = 0.02
    start = time.time()
    reg = cph.registration.registration_generalized_icp(
        source_gpu,
        target_gpu,
        threshold,
        np.eye(4, dtype=np.float32),
        cph.registration.TransformationEstimationForGeneralizedICP(),
        cph.registration.ICPConvergenceCriteria(max_iteration=30),
    )
    print(reg)
    print(reg.transformation)
    print("GICP (GPU) [sec]:", time.time() - start)
    if gt is not None:
        print("distance to the ground-truth pose (Frobenius):", float(np.linalg.norm(reg.transformation - gt)))
