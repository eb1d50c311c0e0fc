"""geometry::OccupancyGrid: ray-cast a scan into a dense log-odds grid (the reference's unit tests
src/tests/geometry/occupancygrid.cpp show the same calls in C++)."""
import numpy as np

import cupoch_b200 as cph

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    d = rng.standard_normal((500_000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    scan = (d * rng.uniform(2.0, 6.0, (len(d), 1))).astype(np.float32)          # a sensor at the origin inside a rough shell
    grid = cph.geometry.OccupancyGrid(voxel_size=0.05, resolution=512)
    grid.insert(cph.geometry.PointCloud(scan), viewpoint=(0.0, 0.0, 0.0), max_range=5.0)
    print(grid)
    occ_idx, occ_prob = grid.extract_occupied_voxels()
    free_idx, free_prob = grid.extract_free_voxels()
    print("occupied voxels:", len(occ_idx), "free voxels:", len(free_idx), "bounds", grid.get_min_bound(), grid.get_max_bound())
    known, voxel = grid.get_voxel((1.0, 0.0, 0.0))
    print("voxel at (1, 0, 0): known", known, voxel, "occupied", grid.is_occupied((1.0, 0.0, 0.0)))
