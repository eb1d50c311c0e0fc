"""Counterpart of the reference's examples/python/advanced/colored_pointcloud_registration.py:37-60: colored ICP over a
3-scale pyramid."""
import time

import numpy as np

from _clouds import pair
import cupoch_b200 as cph

if __name__ == "__main__":
    src, tgt, tc, sc, gt = pair(2_000_000, colors=True, extent=4.0)
    source, target = cph.geometry.PointCloud(src), cph.geometry.PointCloud(tgt)
    source.colors, target.colors = sc, tc
    voxel_radius = [0.05, 0.025, 0.0125]
    max_iter = [50, 30, 14]
    current_transformation = np.identity(4, dtype=np.float32)
    start = time.time()
    for scale in range(3):
        iters, radius = max_iter[scale], voxel_radius[scale]
        print("scale", scale, ": voxel", radius, "iterations", iters)
        source_down = source.voxel_down_sample(radius)
        target_down = target.voxel_down_sample(radius)
        source_down.estimate_normals(cph.geometry.KDTreeSearchParamRadius(radius * 2, 30))
        target_down.estimate_normals(cph.geometry.KDTreeSearchParamRadius(radius * 2, 30))
        result_icp = cph.registration.registration_colored_icp(
            source_down,
            target_down,
            radius,
            current_transformation,
            cph.registration.ICPConvergenceCriteria(relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=iters),
        )
        current_transformation = result_icp.transformation
        print(result_icp)
    print("pyramid [sec]:", time.time() - start)
    print(current_transformation)
    if gt is not None:
        print("distance to the ground-truth pose (Frobenius):", float(np.linalg.norm(current_transformation - gt)))
