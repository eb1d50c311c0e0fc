"""Counterpart of the reference's examples/python/basic/icp_registration.py: point-to-plane ICP."""
import time

import numpy as np

from _clouds import pair
import cupoch_b200 as cph

if __name__ == "__main__":
    src, tgt, _, _, gt = pair()
    source_gpu = cph.geometry.PointCloud(src)
    target_gpu = cph.geometry.PointCloud(tgt)
    threshold = 0.02
    target_gpu.estimate_normals()                     # KDTreeSearchParamKNN(30), like the reference's default
    trans_init = np.eye(4)
    start = time.time()
    reg_p2l = cph.registration.registration_icp(
        source_gpu,
        target_gpu,
        threshold,
        trans_init.astype(np.float32),
        cph.registration.TransformationEstimationPointToPlane(),
    )
    elapsed_time = time.time() - start
    print(reg_p2l)
    print(reg_p2l.transformation)
    print("ICP (GPU) [sec]:", elapsed_time)
    if gt is not None:
        print("distance to the ground-truth pose (Frobenius):", float(np.linalg.norm(reg_p2l.transformation - gt)))
    source_gpu.transform(reg_p2l.transformation)
